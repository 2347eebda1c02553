"""argsort alone on random radii (for ncu launch lists / timing).  python profiles/sort_probe.py [keys] [calls]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyxfocus_b200 import analyses as A  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    g = torch.Generator(device="cuda").manual_seed(0)
    keys = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 1e-3
    for k in range(calls):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        ks, idx = A.argsort(keys)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev = e0.elapsed_time(e1) * 1e-3
        print("call %d: argsort of %d keys: wall %.3f ms, device %.3f ms = %.2f Gkeys/s" % (k, n, wall * 1e3, dev * 1e3,
                                                                                        n / dev / 1e9))
    assert bool((ks[1:] >= ks[:-1]).all())


if __name__ == "__main__":
    main()
