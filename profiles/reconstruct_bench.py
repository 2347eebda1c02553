"""Southwell reconstruction: device pipeline (pxf_reconstruct) vs the CPU restatement of reconstruct.f95, same
input (gradients of a Legendre surface inside a circular aperture, southwell.example()), same sweep count, same
bits.  python profiles/reconstruct_bench.py [sizes...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyxfocus_b200 import reconstruct  # noqa: E402
from oracle import f2py as of  # noqa: E402  (the CPU side of the comparison)


def case(n):
    xg, yg = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-1, 1, n))
    img = np.polynomial.legendre.legval2d(xg, yg, [[0, 1, 0], [0, .5, 0], [1, 0, 0]])
    gx, gy = np.gradient(img)
    out = np.sqrt(xg ** 2 + yg ** 2) > 1
    gx[out] = 100.
    gy[out] = 100.
    ph = np.zeros(gx.shape, order="F")
    ph[out] = 100.

    def pad(a):
        t = np.zeros((a.shape[0] + 2, a.shape[1] + 2), order="F") + 100.
        t[1:-1, 1:-1] = a
        return t
    return [pad(gx), pad(gy), pad(ph)]


def main():
    sizes = [int(v) for v in sys.argv[1:]] or [64, 128, 256, 512]
    print("%6s %8s %12s %12s %8s  %s" % ("grid", "sweeps", "CPU ms", "GPU ms", "ratio", "cell updates/s (GPU)"))
    for n in sizes:
        a = case(n)
        b = [v.copy(order="F") for v in a]
        t0 = time.perf_counter()
        want = of.reconstruct.reconstruct(a[0], a[1], 1e-10, 1., a[2], 100000)
        tc = (time.perf_counter() - t0) * 1e3
        sw = of.reconstruct.reconstruct.sweeps
        reconstruct.reconstruct(*[v.copy(order="F") for v in b[:2]], 1e-10, 1., b[2].copy(order="F"), 100000)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = reconstruct.reconstruct(b[0], b[1], 1e-10, 1., b[2], 100000)
        torch.cuda.synchronize()
        tg = (time.perf_counter() - t0) * 1e3
        assert reconstruct.reconstruct.sweeps == sw and np.array_equal(got, want)
        cells = int((a[2][1:-1, 1:-1] != 100.).sum())
        print("%6d %8d %12.2f %12.2f %8.2f  %.3e" % (n, sw, tc, tg, tc / tg, cells * sw / (tg * 1e-3)))


if __name__ == "__main__":
    main()
