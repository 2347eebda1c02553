"""Segmented (one launch for all shells) vs plain fused programs: kernel time on a resident bundle.
    python profiles/seg_bench.py [rays]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyxfocus_b200 as pxf  # noqa: E402
from pyxfocus_b200 import sources  # noqa: E402
from pyxfocus_b200._call import bundle_alloc  # noqa: E402


def best(fn, reps=7):
    b = 1e30
    for k in range(reps + 2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            b = min(b, e0.elapsed_time(e1))
    return b


def chain(r0, z0):
    return (pxf.Program().transform(0, 0, z0, 0, 0, 0).wolterprimary(r0, z0, 1.).reflect()
            .woltersecondary(r0, z0, 1.).reflect().flat())


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    dev = torch.device("cuda", 0)
    out = bundle_alloc(n, dev)
    src = sources.annulus(220., 220.6, n, zhat=-1., rng="philox", seed=0, device=dev)
    p1 = chain(220., 8400.)
    print("rays %d" % n)
    print("Program (one shell)                    %8.3f ms" % best(lambda: p1.run(src, out=out)))
    for nseg, vary in ((1, False), (260, False), (260, True)):
        per = [2 * (n // nseg // 2)] * nseg
        per[-1] += n - sum(per)
        radii = np.linspace(200., 1500., nseg) if vary else np.full(nseg, 220.)
        z0s = np.sqrt(1.e4 ** 2 - radii ** 2) if vary else np.full(nseg, 8400.)
        if vary:
            src2 = sources.segments("annulus", [(r, r + .6, 0., -1.) for r in radii], per, seed=0, device=dev)
            print("source, %3d segments                   %8.3f ms" % (nseg, best(lambda: sources.segments(
                "annulus", [(r, r + .6, 0., -1.) for r in radii], per, seed=0, out=src2))))
        else:
            src2 = src
        sp = pxf.SegmentedProgram([chain(float(r), float(z)) for r, z in zip(radii, z0s)], per)
        print("SegmentedProgram %3d segments vary=%d    %8.3f ms" % (nseg, vary, best(lambda: sp.run(src2, out=out))))
    if True:
        radii = np.linspace(200., 1500., 260)
        for r in (200., 800., 1500.):
            z0 = float(np.sqrt(1.e4 ** 2 - r ** 2))
            s1 = sources.annulus(r, r + .6, n, zhat=-1., rng="philox", seed=0, device=dev)
            pr = chain(r, z0)
            print("Program, shell r0=%6.1f               %8.3f ms" % (r, best(lambda: pr.run(s1, out=out))))


if __name__ == "__main__":
    main()
