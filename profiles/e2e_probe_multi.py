"""All ranks at once: the bare pinned copies of the end-to-end step's bytes (40 B/ray down || 16 B/ray up per GPU) and
the host-array entry point itself -- the ceiling the N-GPU `e2e` figure of bench.py can be compared with.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/e2e_probe_multi.py [rays/GPU]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyxfocus_b200 as pxf  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def worst(dt):
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t[0])
    host = [None] + [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(9)]
    d = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(7)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for label, down, up in (("40 B/ray down || 16 B/ray up", 5, 2), ("40 B/ray down only", 5, 0), ("16 B/ray up only", 0, 2)):
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s1):
                for k in range(down):
                    host[1 + k].copy_(d[k], non_blocking=True)
            with torch.cuda.stream(s2):
                for k in range(up):
                    d[5 + k].copy_(host[7 + k], non_blocking=True)
            torch.cuda.synchronize()
            dt = worst(time.perf_counter() - t0)
            if rank == 0 and rep > 0:
                print("%d GPUs, raw pinned copies (%s): %.1f ms, %.1f GB/s over all GPUs = a ceiling of %.3e rays/s"
                      % (world, label, dt * 1e3, 8 * (down + up) * n * world / dt / 1e9, n * world / dt), flush=True)
    src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, first=rank * n, device=dev)
    prog = (pxf.Program().transform(0., 0., 8400., 0., 0., 0.).wolterprimary(220., 8400., 1.).reflect()
            .woltersecondary(220., 8400., 1.).reflect().flat())
    pristine = [None] + [src[k].cpu() for k in range(1, 7)]
    for rep in range(3):
        for k in range(1, 7):
            host[k].copy_(pristine[k])
        barrier()
        t0 = time.perf_counter()
        pxf.host.trace(host, prog, write_back=True, hpd=(world == 1), const_rows=pxf.host.SOURCE_CONST_ROWS)
        torch.cuda.synchronize()
        dt = worst(time.perf_counter() - t0)
        if rank == 0 and rep > 0:
            print("%d GPUs, pxf.host.trace (threads/rank: PXF_HOST_THREADS=%s): %.1f ms = %.3e rays/s"
                  % (world, os.environ.get("PXF_HOST_THREADS", "auto"), dt * 1e3, n * world / dt), flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
