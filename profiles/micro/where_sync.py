import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import pyxfocus_b200 as pxf
from pyxfocus_b200 import dist
from pyxfocus_b200._call import bundle_alloc
dev = torch.device("cuda", 0)
n = int(1.25e8)
src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, first=0, device=dev)
out = bundle_alloc(n, dev, zero=True)
prog = (pxf.Program().transform(0., 0., 8400., 0., 0., 0.).wolterprimary(220., 8400., 1.).reflect()
        .woltersecondary(220., 8400., 1.).reflect().flat())
sums = torch.zeros(16, dtype=torch.float64, device=dev)
marks = []
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        t = time.perf_counter(); r = f(self, *a, **k); marks.append((name, t, time.perf_counter())); return r
    setattr(cls, name, g)
for nm in ["__init__", "bracket_params", "small_select", "collect", "cand_hist", "cand_scan", "cand_gather"]:
    wrap(dist.CudaSelect, nm)
for k in range(8):
    marks.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prog.run(src, out=out, sums=sums)
    t1 = time.perf_counter()
    dist.hpd(out, sums=sums, total=n, min_shard=n)
    t2 = time.perf_counter()
print("prog.run host %.3f ms, hpd host %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
for nm, a, b in marks:
    print("%-16s start +%.3f ms  dur %.3f ms" % (nm, (a - t1) * 1e3, (b - a) * 1e3))

# ---- statement-level timing inside collect()
import ctypes
from pyxfocus_b200 import _lib
from pyxfocus_b200._call import stream_ptr
L = _lib.lib()
for k in range(3):
    torch.cuda.synchronize()
    prog.run(src, out=out, sums=sums)
    x, y = out[1], out[2]
    cxy = torch.zeros(2, dtype=torch.float64, device=dev)
    lohi = torch.tensor([0., 1e-6, 2e-6, 1.], dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    prog.run(src, out=out, sums=sums)
    T = [time.perf_counter()]
    cap = int(L.pxf_bracket_capacity(n)); T.append(time.perf_counter())
    cand = torch.empty(cap, dtype=torch.float64, device=dev); T.append(time.perf_counter())
    counters = torch.zeros(5, dtype=torch.int64, device=dev); T.append(time.perf_counter())
    _lib.check(L.pxf_bracket_collect(x.data_ptr(), y.data_ptr(), n, cxy.data_ptr(), lohi.data_ptr(), cand.data_ptr(), cap,
                                     counters.data_ptr(), stream_ptr(dev))); T.append(time.perf_counter())
    counters[3] = (counters[1] > cap).to(torch.int64); T.append(time.perf_counter())
    counters[4] = cap; T.append(time.perf_counter())
    print("capacity %.3f | empty %.3f | zeros %.3f | collect launch %.3f | counters[3]= %.3f | counters[4]= %.3f ms" %
          tuple((T[i + 1] - T[i]) * 1e3 for i in range(6)))
