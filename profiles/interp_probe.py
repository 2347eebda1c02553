"""analyses.interpolateVec on the device against scipy's griddata on the host cores (the reference's own path).
    python profiles/interp_probe.py [points] [grid side]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyxfocus_b200 as pxf  # noqa: E402
from oracle import refapi  # noqa: E402  (the checker: scipy's griddata behind the reference's interpolateVec)


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    side = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    rng = np.random.default_rng(0)
    r, t = 12.5 * np.sqrt(rng.uniform(0, 1, n)), rng.uniform(0, 2 * np.pi, n)
    x, y = r * np.cos(t), r * np.sin(t)
    l = 1e-3 * np.sin(x / 5.) + 1e-6 * rng.normal(size=n)
    zero = np.zeros(n)
    rays = [zero, x, y, zero, l, zero, zero, zero, zero, zero]
    dev = [torch.from_numpy(a).cuda() for a in rays]
    for method in ("linear", "nearest", "cubic"):
        best = 1e30
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got, _, _ = pxf.analyses.interpolateVec(dev, 4, side, side, method=method)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        t0 = time.perf_counter()
        want, _, _ = refapi.interpolateVec(rays, 4, side, side, method=method)
        cpu = time.perf_counter() - t0
        g = got.cpu().numpy()
        same_nan = np.array_equal(np.isnan(g), np.isnan(want))
        err = np.nanmax(np.abs(g - want)) if same_nan else float("nan")
        print("%s: %d points -> %dx%d grid (%.0f%% of it outside the hull): device %.2f ms, scipy (1 core) %.0f ms, x%.0f; "
              "NaN mask equal: %s, max |delta| %.2e" % (method, n, side, side, 100 * np.isnan(want).mean(), best * 1e3, cpu * 1e3,
                                                         cpu / best, same_nan, err), flush=True)


if __name__ == "__main__":
    main()
