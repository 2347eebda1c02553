"""Time prefixes of the Wolter-I chain (fused kernel, out of place, 1.25e8 rays) to see where
the trace kernel's time goes.  Run on the GPU box: python profiles/micro/chain_breakdown.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pyxfocus_b200 as pxf
from pyxfocus_b200._call import bundle_alloc

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
dev = torch.device("cuda", 0)
src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, device=dev)
out = bundle_alloc(n, dev, zero=True)
P = pxf.Program
cases = {
    "transform+flat (6 rows in, 9 out, ~1 div)": P().transform(0, 0, 8400., 0, 0, 0).flat(),
    "transform+primary+reflect": P().transform(0, 0, 8400., 0, 0, 0).wolterprimary(220., 8400., 1.).reflect(),
    "..+secondary+reflect": P().transform(0, 0, 8400., 0, 0, 0).wolterprimary(220., 8400., 1.).reflect()
    .woltersecondary(220., 8400., 1.).reflect(),
    "full chain (+flat)": P().transform(0, 0, 8400., 0, 0, 0).wolterprimary(220., 8400., 1.).reflect()
    .woltersecondary(220., 8400., 1.).reflect().flat(),
}
for name, prog in cases.items():
    for _ in range(3):
        prog.run(src, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        prog.run(src, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-45s %7.3f ms  %6.2f Grays/s" % (name, ms, n / ms * 1e-6), flush=True)
