"""Executed by tests/test_gpu_jit.py in two processes -- PXF_JIT=0 (interpreter only) and PXF_JIT_MIN_RAYS=0 (every
program specialised at run time) -- and prints one line per case: name, sha256 of the resulting rows / flags /
side arrays, and the kernel that ran.  The digests must be identical: the specialised kernels are compositions of
the same per-op device functions."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyxfocus_b200 as pxf  # noqa: E402
from pyxfocus_b200 import examples as ex  # noqa: E402
from util import random_bundle, to_dev  # noqa: E402


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        if t is not None:
            h.update(t.detach().cpu().numpy().tobytes())
    return h.hexdigest()[:24]


def report(name, *tensors):
    print("%s %s %s" % (name, digest(*tensors), pxf.last_trace_kernel().replace(" ", "_")), flush=True)


def main():
    n = 20_011
    P = pxf.Program
    # every opcode, grouped so that each program makes sense on its input
    src = pxf.sources.subannulus(220., 220.6, .4, n, zhat=-1., rng="philox", seed=2, device="cuda")
    wolter = (P().transform(0, 0, 8400., 0, 0, 0).wolterprimary(220., 8400., 1.).reflect().vignette_box(3, 8426., 8526.)
              .vignette_abs(2, 30., 1.5).woltersecondary(220., 8400., 1.).reflect().vignette_mag().flat().vignette_rhogt(3e-6))
    for oop in (False, True):
        rays = [r.clone() for r in src]
        out = [torch.zeros_like(r) for r in src] if oop else None
        sums = torch.zeros(16, dtype=torch.float64, device="cuda")
        alive = wolter.run(rays, out=out, sums=sums)
        report("wolter_vignettes_oop%d" % oop, *(out if oop else rays), alive, sums[:3])
    gen = to_dev(random_bundle(n, seed=5))
    for name, prog in (
            ("transforms_refract", P().transform(1., -2., 3., .1, -.2, .3).itransform(.5, .5, -1., -.3, .1, .2).refract(1., 1.5).reflect()),
            ("conics", P().transform(0, 0, 500., 0, 0, 0).conic(-800., -.5).reflect().conicopd(900., -1., 1.2).flatopd(1.1)),
            ("wolteropd_sine", P().transform(0, 0, 8400., 0, 0, 0).wolterprimaryopd(220., 8400., 1., 1.3).reflect().woltersine(220., 8400., 1e-4, .05)),
            ("spo_kickn_radgrat", P().spocone(700., .0146).kickn(1e-4, -2e-4).reflect().flat().radgrat(2.4e-6, 160. / 11832., -3)),
            ("ws_pair", P().transform(0, 0, 1e4, 0, 0, 0).wsprimary(.005498, 1e4, 1.).kick(.003, 0., -1.).reflect().wssecondary(.005498, 1e4, 1.).reflect().flat())):
        base = src if name in ("wolteropd_sine", "ws_pair") else gen
        if name == "spo_kickn_radgrat":
            base = pxf.sources.subannulus(700., 700.6, .05, n, zhat=-1., rng="philox", seed=3, device="cuda")
        if name == "ws_pair":
            base = pxf.sources.subannulus(220.137, 221.233, .45, n, rng="philox", seed=4, device="cuda")
        rays = [r.clone() for r in base]
        prog.run(rays)
        report(name, *rays)
    # Zernike surface inside a program
    ro, ao = ex.zernike_orders(7)
    z = ex.config3_fast(n, rng="philox", rng_seed=1)
    report("config3", *z["rays"], z["idx"])
    # grating fan + the way back, segmented nested shells, SPO shells
    c4 = ex.config4_fast(40, 72, order=-3, wave=2.4, rng="philox")
    report("config4", *c4["rays"], *c4["surv"])
    c4w = ex.config4_fast(40, 72, order=-1, wave="uniform", rng="philox")
    report("config4_radgratw", *c4w["rays"])
    c5 = ex.config5_fast(77, 260, offaxis=3e-4, rng="philox")
    report("config5", *c5["rays"], c5["weights"])
    print("STATUS " + pxf.jit_status().replace(" ", "_"), flush=True)


if __name__ == "__main__":
    main()
