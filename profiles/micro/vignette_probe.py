import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pyxfocus_b200 as pxf
n = 50_000_000
rays = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0)
flags = rays[1] > 0
for _ in range(3):
    out = pxf.transformations.vignette(rays, ind=flags)
torch.cuda.synchronize()
