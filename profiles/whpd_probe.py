"""Weighted HPD alone on a resident nested-assembly bundle (for ncu launch lists / timing).
    python profiles/whpd_probe.py [rays] [calls]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyxfocus_b200 as pxf  # noqa: E402
from pyxfocus_b200 import analyses as A, sources  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda", 0)
    nshell = 260
    radii = np.linspace(200., 1500., nshell)
    z0s = np.sqrt(1.e4 ** 2 - radii ** 2)
    sizes = [n // nshell] * nshell
    sizes[-1] += n - sum(sizes)
    bundle = sources.segments("annulus", [(r, r + .6, 0., -1.) for r in radii], sizes, seed=0, device=dev)
    pxf.SegmentedProgram([pxf.Program().transform(0, 0, z, 0, 0, 0).wolterprimary(r, z, 1.).reflect()
                          .woltersecondary(r, z, 1.).reflect().flat() for r, z in zip(radii, z0s)], sizes).run(bundle)
    w = torch.repeat_interleave(torch.tensor([2 * np.pi * r * .6 / s for r, s in zip(radii, sizes)], dtype=torch.float64, device=dev),
                                torch.tensor(sizes, device=dev))
    for k in range(calls):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = A.hpd(bundle, weights=w)
        torch.cuda.synchronize()
        print("call %d: weighted hpd %.12e in %.3f ms" % (k, h, (time.perf_counter() - t0) * 1e3))


if __name__ == "__main__":
    main()
