"""Executed under torchrun by tests/test_gpu_dist.py (one rank per GPU, NCCL): the sharded
trace + analyses must reproduce the single-GPU result of the whole bundle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    import pyxfocus_b200 as pxf
    from pyxfocus_b200 import dist

    prog = (pxf.Program().transform(0., 0., 8400., 0., 0., 0.).wolterprimary(220., 8400., 1.).reflect()
            .woltersecondary(220., 8400., 1.).reflect().flat())
    ok = True
    for total in (100_003, 6_000_001):            # five-pass path and bracketed path
        lo, hi = dist.shard_range(total, rank, world)
        shard = pxf.sources.subannulus(220., 220.6, 2 * np.pi, hi - lo, zhat=-1., rng="philox", seed=3, first=lo, device=dev)
        prog.run(shard)
        h = dist.hpd(shard)
        # caller-known sizes: nothing is read back before the result and the centroid sums travel in
        # the sample all-gather (the path bench.py takes); must give the same bits
        h2 = dist.hpd(shard, total=total, min_shard=total // world)
        assert h2 == h, (h2, h)
        rms = dist.rmsCentroid(shard)
        cx, cy = dist.centroid(shard)
        dz = dist.analyticImagePlane(shard)
        # every rank recomputes the whole bundle alone and compares
        whole = pxf.sources.subannulus(220., 220.6, 2 * np.pi, total, zhat=-1., rng="philox", seed=3, first=0, device=dev)
        prog.run(whole)
        assert torch.equal(whole[1][lo:hi], shard[1]) and torch.equal(whole[9][lo:hi], shard[9]), "shard != slice of whole"
        h1 = pxf.analyses.hpd(whole)
        r1 = pxf.analyses.rmsCentroid(whole)
        c1 = pxf.analyses.centroid(whole)
        d1 = pxf.analyses.analyticImagePlane(whole)
        good = (abs(h - h1) <= 1e-9 * abs(h1) and abs(rms - r1) <= 1e-9 * r1 and abs(cx - c1[0]) <= 1e-15
                and abs(cy - c1[1]) <= 1e-15 and abs(dz - d1) <= 1e-6 * max(1., abs(d1)))
        # weighted HPD: small = sorted shards merged by the 64-ary key-space search; large = gathered sample
        # brackets + candidate windows (no rank sorts its shard) -- both vs the single-GPU result
        ww = torch.linspace(.5, 2., total, dtype=torch.float64, device=dev)
        hw = dist.hpd(shard, weights=ww[lo:hi])
        hw1 = pxf.analyses.hpd(whole, weights=ww)
        print("rank %d weighted hpd %.15e vs %.15e" % (rank, hw, hw1), flush=True)
        good = good and abs(hw - hw1) <= 1e-9 * abs(hw1)
        if total > 1_000_000:
            from pyxfocus_b200 import dist as D
            cxw, cyw = D.centroid(shard, ww[lo:hi])
            br = D.CudaWeightedBracket(shard, ww[lo:hi], cxw, cyw)
            W = D.all_reduce_sum(ww[lo:hi].sum())
            res, okb = D.hpd_weighted_bracketed(br, total, W)
            print("rank %d bracketed weighted hpd %.15e valid %s" % (rank, float(res), okb), flush=True)
            good = good and okb and abs(float(res) - hw1) <= 1e-9 * abs(hw1)
        # sharded rhocdf (sample sort, one all-to-all): this rank's slice of the global sorted radii / cdf
        rs, cs, first = dist.rhocdf(shard, weights=ww[lo:hi])
        r1s, c1s = pxf.analyses.rhocdf(whole, weights=ww)
        m = rs.shape[0]
        # (the all-reduced centroid may differ from the single-GPU one in the last bit, and with it a few radii)
        okr = (torch.allclose(rs, r1s[first:first + m], rtol=1e-12, atol=0.)
               and float((cs - c1s[first:first + m]).abs().max()) <= 1e-12)
        cnt = torch.tensor([float(m)], dtype=torch.float64, device=dev)
        td.all_reduce(cnt)
        okr = okr and int(cnt.item()) == total
        print("rank %d rhocdf slice [%d, %d) of %d -> %s" % (rank, first, first + m, total, okr), flush=True)
        good = good and okr
        print("rank %d total %d: hpd %.15e vs %.15e rms %.6e/%.6e dz %.6e/%.6e -> %s" % (rank, total, h, h1, rms, r1, dz, d1, good),
              flush=True)
        ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=dev)
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    td.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
