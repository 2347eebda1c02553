"""BASELINE configs 2-5 end to end on one GPU through the public (reference-shaped) API: source on the
device -> trace (fused where the ops allow) -> vignette -> analyses.  The headline bench (bench.py) is config 1
at config-5 scale; this table is the evidence for the other rows of SURVEY.md 8(d).
    python profiles/config_bench.py [rays] [reps]
Every repetition regenerates the source on the device (Philox, sources.py formulas), so it is a whole pass
of the reference script, not a replay on warm rows.  CUDA-event timed, results read back once per rep (as the
reference scripts do)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyxfocus_b200 as pxf  # noqa: E402
from pyxfocus_b200 import analyses as A, conicsolve, sources, surfaces as S, transformations as T  # noqa: E402
from oracle import chains  # noqa: E402  (geometry constants only: aperture radii, Zernike orders)


def timed(fn, reps):
    """(best, median) CUDA-event time of one pass, wall time of the best pass, last result."""
    fn()
    fn()
    ts, walls, out = [], [], None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - t0) * 1e3)
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts)), min(walls), out


STAGES = os.environ.get("PXF_STAGES") == "1"


class stage:
    """with stage("name"): ...  prints the stage's GPU time when PXF_STAGES=1 (adds a sync per stage)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if STAGES:
            torch.cuda.synchronize()
            self.t0 = time.perf_counter()

    def __exit__(self, *a):
        if STAGES:
            torch.cuda.synchronize()
            print("    %-28s %8.3f ms" % (self.name, (time.perf_counter() - self.t0) * 1e3))
        return False


def config1(n, dev):
    """BASELINE configs[0] at its own size (1e5 rays, examples/axro/singlePassAlignment.py:246-269): numpy-seeded
    source (host MT19937 draws, uploaded), fused trace, hpd.  Launch-latency territory."""
    n = 100_000

    def step():
        np.random.seed(0)
        with stage("source (numpy draws + upload)"):
            rays = sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., device=dev)
        with stage("trace (fused, 6 ops)"):
            with pxf.fused(rays):
                T.transform(rays, 0, 0, -8400., 0, 0, 0)
                S.wolterprimary(rays, 220., 8400.)
                T.reflect(rays)
                S.woltersecondary(rays, 220., 8400.)
                T.reflect(rays)
                S.flat(rays)
        with stage("hpd"):
            res = dict(hpd=A.hpd(rays))
        return res
    return step, 3, n


def config2(n, dev):
    """W-S shell, one off-axis field point (5 arcmin), focusI + findimageplane x2 + hpd + rms
    (examples/axro/axialHeights.py:77-113, WSverify.py:159-167)."""
    r0, z0, psi = 220., 1.e4, 1.
    a0, a1 = chains.ws_aperture(r0, z0, psi, 200.)
    theta = 5. / 60. * np.pi / 180.
    alpha = conicsolve.woltparam(r0, z0)[0]

    def step():
        with stage("source"):
            rays = sources.subannulus(a0, a1, 100. / r0, n, zhat=-1., rng="philox", seed=0, device=dev)
        # rays[4] += sin(theta); rays[6] = -sqrt(1-rays[4]**2) of the reference script is the program's `kick`
        with stage("trace (fused, 6 ops)"):
            (pxf.Program().transform(0, 0, z0, 0, 0, 0).wsprimary(alpha, z0, psi).kick(np.sin(theta), 0., -1.).reflect()
             .wssecondary(alpha, z0, psi).reflect()).run(rays)
        with stage("focusI"):
            dz = S.focusI(rays)
        with stage("findimageplane x2"):
            d1 = A.findimageplane(rays, 20., 100)
            d2 = A.findimageplane(rays, 1., 100)
        with stage("hpd + rms"):
            res = dict(dz=dz, scan=(d1, d2), hpd=A.hpd(rays), rms=A.rmsCentroid(rays))
        return res
    return step, 2


def config3(n, dev):
    """Zernike-perturbed wavefront (36 terms) through a Wolter-I pair with two vignettes
    (examples/axro/singlePassAlignment.py:22-56,133-187; slf.py:145-147)."""
    ro, ao = chains.zernike_orders(7)
    coeff = chains.zernike_coeff(len(ro), seed=0, sigma=1.e-4)

    def step():
        with stage("source"):
            rays = sources.subannulus(220., 220.6, 100. / 220., n, zhat=-1., rng="philox", seed=0, device=dev)
        # one fused launch: transform -> zernsurf (table in shared memory) -> the Wolter-I pair with its vignettes
        with stage("fused trace (14 ops incl. the Zernike surface)"):
            prog = (pxf.Program().transform(-220.3, 0, 100., 0, 0, 0).zernsurf(coeff, ro, ao, 62.5, 1.)
                    .reflect().transform(0, 0, 0, -np.pi, 0, 0).flatopd(1.)
                    .transform(220.3, 0, 8600., 0, 0, 0)
                    .wolterprimary(220., 8400., 1.).reflect()
                    .vignette_box(3, 8426., 8526.).vignette_abs(2, 50.)
                    .woltersecondary(220., 8400., 1.).reflect().vignette_mag().flat())
            alive = prog.run(rays)
        with stage("vignette (compaction)"):
            surv = T.compact(rays, alive)      # alive: uint8 flags from the program's vignette predicates
        with stage("hpd + rms"):
            res = dict(kept=surv[1].shape[0] / n, hpd=A.hpd(surv), rms=A.rmsCentroid(surv))
        return res
    return step, 4


def config4(n, dev):
    """SPO pair + radial grating, one diffraction order with a scalar wavelength on half the rays and a
    per-ray wavelength (radgratW) on the other half (examples/arcus/cat.py:203-288, slf.py:61-167)."""
    R0, F, hub = 700., 12.e3, 11832.911
    half = torch.zeros(n, dtype=torch.bool, device=dev)
    half[: n // 2] = True
    wave = torch.empty(n, dtype=torch.float64, device=dev).uniform_(3.6e-6, 7.2e-6)   # full length: wave[ind]

    def step():
        with stage("source"):
            rays = sources.subannulus(R0, R0 + .605, .05, n, zhat=-1., rng="philox", seed=0, device=dev)
        stg = stage("SPO pair (fused) + masked grating")
        stg.__enter__()
        with pxf.fused(rays):
            T.transform(rays, 0, 0, 0, 0, 0, .01)
            S.spoPrimary(rays, R0, F)
            T.reflect(rays)
            S.spoSecondary(rays, R0, F)
            T.reflect(rays)
            T.transform(rays, 0, 0, -(F - 200.), 0, 0, 0)
            S.flat(rays)
            T.transform(rays, 0, hub, 0, 0, 0, 0)
        T.reflect(rays, ind=half)
        T.radgrat(rays, 160. / hub, -3, 2.4e-6, ind=half)
        T.radgrat(rays, 160. / hub, 1, wave, ind=~half)
        with pxf.fused(rays):
            T.transform(rays, 0, -hub, 0, 0, 0, 0)
            T.transform(rays, 0, 0, -200., 0, 0, 0)
            S.flat(rays)
        stg.__exit__()
        with stage("vignette (compaction)"):
            surv = T.vignette(rays)
        with stage("centroid + rmsY"):
            res = dict(kept=surv[1].shape[0] / n, cent=A.centroid(surv), rmsY=A.rmsY(surv))
        return res
    return step, 3


def config5(n, dev, nshell=260):
    """Nested Wolter-I assembly: one prescription per shell, one fused launch per shell segment, area
    weights, weighted centroid / rms / hpd (examples/axro/axialHeights.py:215-322, SMARTX.py:163-259)."""
    from pyxfocus_b200._call import bundle_alloc, bundle_split
    radii = np.linspace(200., 1500., nshell)
    per = [2 * (n // nshell // 2)] * nshell
    total = sum(per)
    w = torch.empty(total, dtype=torch.float64, device=dev)
    off = 0
    for r0, nk in zip(radii, per):
        w[off:off + nk] = 2 * np.pi * r0 * .6 / nk
        off += nk
    bundle = bundle_alloc(total, dev, zero=True)
    segs = bundle_split(bundle, per)

    z0s = [float(np.sqrt(1.e4 ** 2 - r0 ** 2)) for r0 in radii]
    seg = pxf.SegmentedProgram([pxf.Program().transform(0, 0, z0, 0, 0, 0).wolterprimary(r0, z0, 1.).reflect()
                                .woltersecondary(r0, z0, 1.).reflect().flat() for r0, z0 in zip(radii, z0s)], per)

    def step_segmented():
        with stage("source + trace (2 launches)"):
            sources.segments("annulus", [(r0, r0 + .6, 0., -1.) for r0 in radii], per, seed=0, out=bundle)
            seg.run(bundle)
        with stage("weighted hpd"):
            h = A.hpd(bundle, weights=w)
        with stage("weighted rms + centroid"):
            res = dict(hpd_w=h, rms_w=A.rmsCentroid(bundle, weights=w), cent=A.centroid(bundle, weights=w))
        return res
    if os.environ.get("PXF_PER_SHELL") != "1":
        return step_segmented, 3, total

    def step():
        off = 0
        with stage("sources + traces (per shell)"):
            for k, (r0, nk) in enumerate(zip(radii, per)):
                z0 = float(np.sqrt(1.e4 ** 2 - r0 ** 2))
                sources.annulus(r0, r0 + .6, nk, zhat=-1., rng="philox", seed=0, first=off, out=segs[k])
                (pxf.Program().transform(0, 0, z0, 0, 0, 0).wolterprimary(r0, z0, 1.).reflect()
                 .woltersecondary(r0, z0, 1.).reflect().flat()).run(segs[k])
                off += nk
        with stage("weighted hpd"):
            h = A.hpd(bundle, weights=w)
        with stage("weighted rms + centroid"):
            res = dict(hpd_w=h, rms_w=A.rmsCentroid(bundle, weights=w), cent=A.centroid(bundle, weights=w))
        return res
    return step, 3, total


def program_only(n, dev, reps):
    """The generic fused interpreter alone (config 3's second program, out of place, rows resident)."""
    from pyxfocus_b200._call import bundle_alloc
    ro, ao = chains.zernike_orders(7)
    coeff = chains.zernike_coeff(len(ro), seed=0, sigma=1.e-4)
    rays = sources.subannulus(220., 220.6, 100. / 220., n, zhat=-1., rng="philox", seed=0, device=dev)
    T.transform(rays, 220.3, 0, -100., 0, 0, 0)
    S.zernsurf(rays, coeff, 62.5, rorder=ro, aorder=ao, nr=1.)
    out = bundle_alloc(n, dev)
    alive = torch.empty(n, dtype=torch.uint8, device=dev)
    prog = (pxf.Program().reflect().transform(0, 0, 0, -np.pi, 0, 0).flatopd(1.)
            .transform(220.3, 0, 8600., 0, 0, 0)
            .wolterprimary(220., 8400., 1.).reflect()
            .vignette_box(3, 8426., 8526.).vignette_abs(2, 50.)
            .woltersecondary(220., 8400., 1.).reflect().vignette_mag().flat())
    best = 1e30
    for k in range(reps + 3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prog.run(rays, alive=alive, out=out)
        e1.record()
        torch.cuda.synchronize()
        if k >= 3:
            best = min(best, e0.elapsed_time(e1))
    print("k_program (config 3 tail, 12 ops, %d rays): best %.3f ms = %.2f Grays/s, kept %.3f" %
          (n, best, n / best / 1e6, float(alive.float().mean())))


def main():
    if len(sys.argv) > 3 and sys.argv[3] == "prog":
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        return program_only(int(float(sys.argv[1])), dev, int(sys.argv[2]))
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    which = sys.argv[3].split(",") if len(sys.argv) > 3 else ["1", "2", "3", "4", "5"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    print("rays per pass: %d, reps %d" % (n, reps))
    print("%-8s %10s %10s %10s %12s %14s  %s" % ("config", "ms (best)", "ms (med)", "ms (wall)", "Grays/s", "Ginteract/s", "result"))
    for name, make in (("1", config1), ("2", config2), ("3", config3), ("4", config4), ("5", config5)):
        if name not in which:
            continue
        made = make(n, dev)
        step, nsurf = made[0], made[1]
        nn = made[2] if len(made) > 2 else n
        l0 = pxf.launch_count()
        ms, med, wall, out = timed(step, reps)
        launches = (pxf.launch_count() - l0) / (reps + 2)
        res = ", ".join("%s=%s" % (k, ("%.6g" % v) if isinstance(v, float) else
                                   "(" + ", ".join("%.6g" % q for q in v) + ")") for k, v in out.items())
        print("%-8s %10.3f %10.3f %10.3f %12.3f %14.3f  %s  [%d launches/pass]" %
              (name, ms, med, wall, nn / ms / 1e6, nn * nsurf / ms / 1e6, res, launches))


if __name__ == "__main__":
    main()
