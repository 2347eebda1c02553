"""Host-array entry point alone (pxf_host_trace_program): PCIe-bound end-to-end step with the library's own
stage timings (PXF_HOST_DEBUG=1) and a plain pinned D2H/H2D copy of the same bytes for comparison.
    PXF_HOST_DEBUG=1 python profiles/e2e_probe.py [rays]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyxfocus_b200 as pxf  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
    dev = torch.device("cuda", 0)
    src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, device=dev)
    prog = (pxf.Program().transform(0., 0., 8400., 0., 0., 0.).wolterprimary(220., 8400., 1.).reflect()
            .woltersecondary(220., 8400., 1.).reflect().flat())
    host = [None] + [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(9)]
    pristine = [None] + [src[k].cpu() for k in range(1, 7)]
    for rep in range(3):
        for k in range(1, 7):
            host[k].copy_(pristine[k])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = pxf.host.trace(host, prog, write_back=True, hpd=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("rep %d: %.1f ms = %.3e rays/s, hpd %.6e" % (rep, dt * 1e3, n / dt, r["hpd"]), flush=True)
    # raw PCIe: 5 rows down, 2 rows up, on two streams at once
    d = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(7)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            for k in range(5):
                host[1 + k].copy_(d[k], non_blocking=True)
        with torch.cuda.stream(s2):
            for k in range(2):
                d[5 + k].copy_(host[7 + k], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("raw copies (40 B/ray down || 16 B/ray up): %.1f ms = %.1f GB/s down" % (dt * 1e3, 40 * n / dt / 1e9), flush=True)


if __name__ == "__main__":
    main()
