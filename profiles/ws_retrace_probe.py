"""Cost of PXF_OPT_WS_RETRACE (re-trace long-trip / restored W-S rays with the exact form) on- and off-axis."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyxfocus_b200 as pxf  # noqa: E402
from pyxfocus_b200 import transformationsf as TF, woltsurf as WS  # noqa: E402
from pyxfocus_b200._call import bundle_alloc  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
dev = torch.device("cuda", 0)
alpha = pxf.conicsolve.woltparam(220., 1e4)[0]
for arcmin in (0., 10., 24.):
    th = arcmin / 60. * np.pi / 180.
    for retrace in (0, 12):
        pxf.set_option(pxf.OPT_WS_RETRACE, retrace)
        best = [1e30, 1e30]
        for rep in range(3):
            r = pxf.sources.subannulus(220.13, 221.23, 100. / 220., n, zhat=-1., rng="philox", seed=0, device=dev)
            TF.transform(*r[1:], 0., 0., -1e4, 0., 0., 0.)
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record(); WS.wsprimary(*r[1:], alpha, 1e4, 1.); e[1].record()
            r[4].add_(np.sin(th)); r[6].copy_(-(1. - r[4] ** 2).sqrt())
            TF.reflect(*r[4:])
            e[2].record(); WS.wssecondary(*r[1:], alpha, 1e4, 1.); e[3].record()
            torch.cuda.synchronize()
            best = [min(best[0], e[0].elapsed_time(e[1])), min(best[1], e[2].elapsed_time(e[3]))]
        print("%4.0f' retrace=%2d: wsprimary %.3f ms, wssecondary %.3f ms per %d rays" % (arcmin, retrace, best[0], best[1], n))
pxf.set_option(pxf.OPT_WS_RETRACE, 0)
