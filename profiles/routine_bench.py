"""Per-routine throughput against the HBM roofline (SURVEY.md 8a "B/ray" column): every f2py-replacement
kernel alone on a resident bundle, CUDA-event timed, best of 5 after warm-up.
    python profiles/routine_bench.py [rays]   ->  table on stdout
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyxfocus_b200 as pxf  # noqa: E402
from pyxfocus_b200 import surfacesf as SF, transformationsf as TF, woltsurf as WS, zernsurf as ZS  # noqa: E402
from pyxfocus_b200._call import bundle_alloc  # noqa: E402


def timed_b2b(fn, calls=20):
    """Average of `calls` back-to-back invocations (each reads its scalar back): keeps the GPU busy, so
    sub-millisecond analyses are not timed at idle clocks."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / calls


def timed(fn, reset, reps=5):
    best = 1e30
    for k in range(reps + 2):
        reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if k >= 2:
            best = min(best, e0.elapsed_time(e1))
    return best


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
    dev = torch.device("cuda", 0)
    peak = 6555.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, device=dev)
    # states along the Wolter-I chain, kept as pristine copies to reset from
    st0 = bundle_alloc(n, dev)
    st1 = bundle_alloc(n, dev)      # after transform (above the primary)
    st2 = bundle_alloc(n, dev)      # after primary + reflect (above the secondary)
    st3 = bundle_alloc(n, dev)      # after secondary + reflect (above the focal plane)
    work = bundle_alloc(n, dev)

    def cp(dst, s_):
        for a, b in zip(dst, s_):
            a.copy_(b)
    cp(st0, src)
    cp(work, src)
    TF.transform(*work[1:], 0., 0., 8400., 0., 0., 0.)
    cp(st1, work)
    WS.wolterprimary(*work[1:], 220., 8400., 1.)
    TF.reflect(*work[4:])
    cp(st2, work)
    WS.woltersecondary(*work[1:], 220., 8400., 1.)
    TF.reflect(*work[4:])
    cp(st3, work)
    rows = []

    def add(name, ref, bytes_per_ray, fn, state):
        ms = timed(fn, lambda: cp(work, state))
        gbs = bytes_per_ray * n / (ms * 1e-3) / 1e9
        rows.append((name, ref, bytes_per_ray, ms, n / (ms * 1e-3), gbs, gbs / peak))

    W = work
    add("transform (general angles)", "transformationsf.f95:134-163", 144, lambda: TF.transform(*W[1:], 1., 2., 3., .1, .2, .3), st1)
    add("transform (translation only)", "transformationsf.f95:134-163", 144, lambda: TF.transform(*W[1:], 0., 0., 8400., 0., 0., 0.), st0)
    add("itransform", "transformationsf.f95:168-201", 144, lambda: TF.itransform(*W[1:], 1., 2., 3., .1, .2, .3), st1)
    add("reflect", "transformationsf.f95:60-79", 72, lambda: TF.reflect(*W[4:]), st2)
    add("refract", "transformationsf.f95:82-130", 96, lambda: TF.refract(*W[4:], 1., 1.5), st2)
    add("flat", "surfacesf.f95:4-29", 96, lambda: SF.flat(*W[1:]), st3)
    add("flatopd", "surfacesf.f95:32-53", 112, lambda: SF.flatopd(*W[1:], W[0], 1.), st3)
    add("wolterprimary", "woltsurf.f95:7-54", 96, lambda: WS.wolterprimary(*W[1:], 220., 8400., 1.), st1)
    add("woltersecondary", "woltsurf.f95:114-161", 96, lambda: WS.woltersecondary(*W[1:], 220., 8400., 1.), st2)
    add("woltersine", "woltsurf.f95:167-215", 96, lambda: WS.woltersine(*W[1:], 220., 8400., 1e-4, .05), st1)
    alpha = pxf.conicsolve.woltparam(220., 8400.)[0]
    add("wsprimary", "woltsurf.f95:387-476", 96, lambda: WS.wsprimary(*W[1:], alpha, 8400., 1.), st1)
    add("conic (paraboloid)", "surfacesf.f95:302-360", 120, lambda: SF.conic(*W[1:], 2e4, -1.), st3)
    add("tracesphere", "surfacesf.f95:57-101", 120, lambda: SF.tracesphere(*W[1:], 9000.), st3)
    add("spocone", "woltsurf.f95:591-638", 96, lambda: WS.spocone(*W[1:], 220., .0065), st1)
    add("radgrat", "transformationsf.f95:205-238", 64, lambda: TF.radgrat(W[1], W[2], W[4], W[5], W[6], 2.4e-6, 160. / 11832.911, -1.), st3)
    ro = [r for r in range(8) for _ in range(r + 1)]
    ao = [m for r in range(8) for m in range(-r, r + 1, 2)]
    zc = np.random.default_rng(0).normal(0., 1e-4, len(ro))
    zc[:3] = 0.
    add("tracezern (36 terms, n<=7)", "zernsurf.f95:8-101", 96, lambda: ZS.tracezern(*W[1:], zc, np.array(ro), np.array(ao), 230.), st3)
    # Legendre-Legendre figure error on the Wolter primary (the reference's figure-error model, woltsurf.f95:219-288):
    # all axial x azimuthal orders up to 5 (36 coefficients)
    llc = np.random.default_rng(1).normal(0., 1e-5, 36)
    lla = np.repeat(np.arange(6), 6).astype(np.int32)
    llz = np.tile(np.arange(6), 6).astype(np.int32)
    add("wolterprimll (36 LL terms, orders<=5)", "woltsurf.f95:219-288", 96,
        lambda: WS.wolterprimll(*W[1:], 220., 8400., 8500., 8400., 2 * np.pi, llc, lla, llz), st1)
    # analyses / compaction
    x, y = st3[1], st3[2]
    rows.append(("centroid", "analyses.py:16-22", 16, *(lambda ms: (ms, n / (ms * 1e-3), 16 * n / (ms * 1e-3) / 1e9, 16 * n / (ms * 1e-3) / 1e9 / peak))(
        timed_b2b(lambda: pxf.analyses.centroid(st3)))))
    rows.append(("rmsCentroid", "analyses.py:24-30", 32, *(lambda ms: (ms, n / (ms * 1e-3), 32 * n / (ms * 1e-3) / 1e9, 32 * n / (ms * 1e-3) / 1e9 / peak))(
        timed_b2b(lambda: pxf.analyses.rmsCentroid(st3)))))
    rows.append(("hpd (unweighted, incl. centroid pass)", "analyses.py:88-97", 32, *(lambda ms: (ms, n / (ms * 1e-3), 32 * n / (ms * 1e-3) / 1e9, 32 * n / (ms * 1e-3) / 1e9 / peak))(
        timed_b2b(lambda: pxf.analyses.hpd(st3)))))
    rows.append(("analyticImagePlane", "analyses.py:118-133", 40, *(lambda ms: (ms, n / (ms * 1e-3), 40 * n / (ms * 1e-3) / 1e9, 40 * n / (ms * 1e-3) / 1e9 / peak))(
        timed_b2b(lambda: pxf.analyses.analyticImagePlane(st3)))))
    wts = torch.linspace(.5, 2., n, dtype=torch.float64, device=dev)
    rows.append(("hpd (weighted)", "analyses.py:88-97", 40, *(lambda ms: (ms, n / (ms * 1e-3), 40 * n / (ms * 1e-3) / 1e9, 40 * n / (ms * 1e-3) / 1e9 / peak))(
        timed_b2b(lambda: pxf.analyses.hpd(st3, weights=wts), calls=5))))
    flags = (st3[1] > 0)
    f = float(flags.float().mean())
    b = 80 + 80 * f + 1
    rows.append(("vignette (mask, %.0f%% kept)" % (100 * f), "transformations.py:214-225", round(b, 1),
                 *(lambda ms: (ms, n / (ms * 1e-3), b * n / (ms * 1e-3) / 1e9, b * n / (ms * 1e-3) / 1e9 / peak))(
                     timed_b2b(lambda: pxf.transformations.vignette(st3, ind=flags), calls=10))))
    rad = pxf.analyses.rho(st3, cent=True)
    # algorithmic traffic of an argsort: keys read once, sorted keys and the int64 permutation written once
    rows.append(("argsort (stable LSD radix, 64-bit keys)", "analyses.py:76", 24,
                 *(lambda ms: (ms, n / (ms * 1e-3), 24 * n / (ms * 1e-3) / 1e9, 24 * n / (ms * 1e-3) / 1e9 / peak))(
                     timed_b2b(lambda: pxf.analyses.argsort(rad), calls=5))))
    # the step either side of the path (round 2)
    def row(name, ref, b, ms):
        rows.append((name, ref, b, ms, n / (ms * 1e-3), b * n / (ms * 1e-3) / 1e9, b * n / (ms * 1e-3) / 1e9 / peak))
    A, T, S = pxf.analyses, pxf.transformations, pxf.sources
    tmp = bundle_alloc(n, dev)
    row("rectbeam (Philox)", "sources.py:348-379", 80, timed_b2b(lambda: S.rectbeam(12., 7., n, rng="philox", out=tmp), calls=10))
    row("convergingbeam (Philox)", "sources.py:250-296", 80,
        timed_b2b(lambda: S.convergingbeam(8400., 200., 230., -.1, .3, n, 1.5, rng="philox", out=tmp), calls=10))
    row("gaussianBeam (Philox + Box-Muller)", "sources.py:381-416", 80, timed_b2b(lambda: S.gaussianBeam(.01, n, rng="philox", out=tmp), calls=10))
    side = int(np.sqrt(n))
    grid_rows = bundle_alloc(side * side, dev)
    ms = timed_b2b(lambda: S.rectArray(4., 2.5, side, out=grid_rows), calls=10)
    rows.append(("rectArray (%d^2)" % side, "sources.py:210-247", 80, ms, side * side / (ms * 1e-3), 80 * side * side / (ms * 1e-3) / 1e9,
                 80 * side * side / (ms * 1e-3) / 1e9 / peak))
    for k in range(10):
        tmp[k].copy_(st3[k])
    row("pointTo", "transformations.py:91-100", 48, timed_b2b(lambda: T.pointTo(tmp, 0., 0., -100.), calls=10))
    row("indAngle", "analyses.py:164-182", 56, timed_b2b(lambda: A.indAngle(st3), calls=10))
    row("measureOPD", "analyses.py:232-244", 32, timed_b2b(lambda: A.measureOPD(st3, (0., 0., 1.)), calls=10))
    row("rmsPoint", "analyses.py:33-45", 24, timed_b2b(lambda: A.rmsPoint(st3, (0., 0., 1.)), calls=10))
    del tmp, grid_rows
    print("rays per launch: %d; HBM peak (measured copy rate): %.1f GB/s" % (n, peak))
    print("%-40s %-30s %7s %9s %11s %9s %6s" % ("routine", "reference", "B/ray", "ms", "Grays/s", "GB/s", "frac"))
    for r in rows:
        print("%-40s %-30s %7s %9.3f %11.2f %9.0f %6.2f" % (r[0], r[1], r[2], r[3], r[4] / 1e9, r[5], r[6]))


if __name__ == "__main__":
    main()
