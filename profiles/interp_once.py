"""One linear and one cubic interpolateVec call (1e6 points -> 512 x 512), for ncu launch lists."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyxfocus_b200 as pxf  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
r, t = 12.5 * np.sqrt(rng.uniform(0, 1, n)), rng.uniform(0, 2 * np.pi, n)
x, y = r * np.cos(t), r * np.sin(t)
l = 1e-3 * np.sin(x / 5.) + 1e-6 * rng.normal(size=n)
z = np.zeros(n)
dev = [torch.from_numpy(a).cuda() for a in [z, x, y, z, l, z, z, z, z, z]]
for method in ("linear", "cubic"):
    got, _, _ = pxf.analyses.interpolateVec(dev, 4, 512, 512, method=method)
    torch.cuda.synchronize()
    print(method, float(torch.nansum(got)))
