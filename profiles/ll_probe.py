import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import pyxfocus_b200 as pxf
from pyxfocus_b200 import woltsurf as WS, transformationsf as TF
from pyxfocus_b200._call import bundle_alloc
n = 50_000_000
dev = torch.device("cuda", 0)
src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, device=dev)
TF.transform(*src[1:], 0., 0., 8400., 0., 0., 0.)
W = bundle_alloc(n, dev)
def run(fn):
    best = 1e9
    for k in range(4):
        for a, b in zip(W, src): a.copy_(b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if k: best = min(best, e0.elapsed_time(e1))
    return best
for orders in (0, 1, 3, 5, 7):
    k = orders + 1
    c = np.random.default_rng(1).normal(0., 1e-5, k * k)
    ax = np.repeat(np.arange(k), k).astype(np.int32)
    az = np.tile(np.arange(k), k).astype(np.int32)
    t = run(lambda: WS.wolterprimll(*W[1:], 220., 8400., 8500., 8400., 2 * np.pi, c, ax, az))
    print("wolterprimll orders<=%d (%2d terms): %.3f ms" % (orders, k * k, t))
print("wolterprimary: %.3f ms" % run(lambda: WS.wolterprimary(*W[1:], 220., 8400., 1.)))
