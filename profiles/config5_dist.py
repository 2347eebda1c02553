"""BASELINE config 5 proper on N GPUs (torchrun, one rank per GPU): nested 260-shell Wolter-I assembly, rays
sharded contiguously across ranks (a shard cuts through shells), area weights, weighted centroid / rms / HPD
over NCCL.  Every pass regenerates the source on the device.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \
        profiles/config5_dist.py [rays per GPU] [reps]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    import pyxfocus_b200 as pxf
    from pyxfocus_b200 import dist, sources
    from pyxfocus_b200._call import bundle_alloc

    per_gpu = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    nshell = 260
    total = per_gpu * world
    radii = np.linspace(200., 1500., nshell)
    z0s = np.sqrt(1.e4 ** 2 - radii ** 2)
    shell_n = np.full(nshell, total // nshell, dtype=np.int64)
    shell_n[-1] += total - shell_n.sum()
    gstart = np.concatenate([[0], np.cumsum(shell_n)])
    lo, hi = dist.shard_range(total, rank, world)
    # this rank's pieces of the shells
    sizes, params, progs, wts = [], [], [], []
    for k in range(nshell):
        a, b = max(lo, gstart[k]), min(hi, gstart[k + 1])
        if b <= a:
            continue
        sizes.append(int(b - a))
        params.append((radii[k], radii[k] + .6, 0., -1.))
        progs.append(pxf.Program().transform(0, 0, z0s[k], 0, 0, 0).wolterprimary(radii[k], z0s[k], 1.).reflect()
                     .woltersecondary(radii[k], z0s[k], 1.).reflect().flat())
        wts.append(2 * np.pi * radii[k] * .6 / shell_n[k])
    seg = pxf.SegmentedProgram(progs, sizes, device=dev)
    w = torch.repeat_interleave(torch.tensor(wts, dtype=torch.float64, device=dev), torch.tensor(sizes, device=dev))
    bundle = bundle_alloc(hi - lo, dev, zero=True)

    def step():
        sources.segments("annulus", params, sizes, seed=0, first=lo, out=bundle)
        seg.run(bundle)
        h = dist.hpd(bundle, weights=w)
        return h, dist.rmsCentroid(bundle, weights=w), dist.centroid(bundle, weights=w)

    for _ in range(2):
        out = step()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        ts.append(float(t.item()))
    # stage split of one more pass (host-synchronised)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sources.segments("annulus", params, sizes, seed=0, first=lo, out=bundle)
    seg.run(bundle)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dist.hpd(bundle, weights=w)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if rank == 0:
        ms = min(ts)
        print("config 5: %d GPUs, %d rays (%d per GPU), %d shells: best %.3f ms median %.3f ms per pass = %.3e rays/s; "
              "hpd_w %.9e rms_w %.6e; source+trace %.3f ms, weighted hpd %.3f ms"
              % (world, total, hi - lo, nshell, ms, float(np.median(ts)), total / ms * 1e3, out[0], out[1],
                 (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
