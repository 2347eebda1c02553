import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyxfocus_b200 as pxf
n = 200000
rng = np.random.default_rng(0)
r, t = 12.5 * np.sqrt(rng.uniform(0, 1, n)), rng.uniform(0, 2 * np.pi, n)
x, y = r * np.cos(t), r * np.sin(t)
l = 1e-3 * np.sin(x / 5.) + 1e-6 * rng.normal(size=n)
z = np.zeros(n)
dev = [torch.from_numpy(a).cuda() for a in [z, x, y, z, l, z, z, z, z, z]]
got, _, _ = pxf.analyses.interpolateVec(dev, 4, 256, 256, method="cubic")
torch.cuda.synchronize()
print(float(torch.nansum(got)))
