"""Where does the multi-GPU HPD spend its time?  torchrun --nproc-per-node 2 profiles/micro/dist_hpd_timing.py
Prints per rank: GPU time of the trace kernel, GPU time from the end of the trace to the HPD result, and the
host time spent enqueueing the HPD (before the final read-back)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    import pyxfocus_b200 as pxf
    from pyxfocus_b200 import dist
    from pyxfocus_b200._call import bundle_alloc
    n = int(float(os.environ.get("RAYS", "1.25e8")))
    src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, first=rank * n, device=dev)
    out = bundle_alloc(n, dev, zero=True)
    prog = (pxf.Program().transform(0., 0., 8400., 0., 0., 0.).wolterprimary(220., 8400., 1.).reflect()
            .woltersecondary(220., 8400., 1.).reflect().flat())
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    orig_small = dist.CudaSelect.small_select
    host_enq = []

    def patched(self, *a, **k):
        if k.get("read"):
            host_enq.append(time.perf_counter())
        return orig_small(self, *a, **k)
    dist.CudaSelect.small_select = patched
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(40)]
    t_host = []
    for k in range(40):
        ev[k][0].record()
        prog.run(src, out=out, sums=sums)
        ev[k][1].record()
        t0 = time.perf_counter()
        dist.hpd(out, sums=sums, total=n * world, min_shard=n)
        ev[k][2].record()
        t_host.append((host_enq[-1] - t0, time.perf_counter() - t0))
    torch.cuda.synchronize()
    tr = np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(10, 40)])
    hp = np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(10, 40)])
    print("rank %d: trace %.3f ms, trace-end -> hpd result %.3f ms (GPU events); host: enqueue hpd %.3f ms, "
          "whole hpd call %.3f ms" % (rank, tr, hp, 1e3 * np.mean([a for a, b in t_host[10:]]),
                                       1e3 * np.mean([b for a, b in t_host[10:]])), flush=True)
    if world > 1:
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
