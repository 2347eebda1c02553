"""Generate the golden fixtures in this directory.

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py
    python tests/golden/make_golden.py --fortran-source     # re-derive every fixture from the reference's Fortran TEXT
                                                            # (oracle/f95run.py in the f2py slots) and compare, writing nothing

It imports the reference's *unmodified* Python layer (sources / transformations / surfaces /
analyses) through ``oracle.refload`` -- with the C oracle standing in for the four f2py
Fortran modules, which cannot be compiled here (no Fortran compiler) -- runs the BASELINE
configurations at small sizes and stores inputs and outputs as .npz.  The GPU parity tests
upload the stored inputs, run the CUDA engine and compare with the stored outputs; the CPU
tests re-run the oracle restatement (oracle/pyref.py, oracle/chains.py) against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import chains, refload  # noqa: E402


def copy(rays):
    return [np.array(r, dtype=np.float64, copy=True) for r in rays]


def pack(prefix, rays):
    return {prefix: np.stack([np.asarray(r, dtype=np.float64) for r in rays])}


def main():
    fortran = "--fortran-source" in sys.argv
    if fortran:
        # the unmodified Python layer over the unmodified Fortran source text: the whole reference.  Nothing is written;
        # every array the script would store is compared with the committed fixture instead.
        from oracle import f95mods
        ref = refload.load(f2py_modules=f95mods.modules())
        report = []

        def compare(path, **g):
            old = np.load(path)
            bad = [k for k in g if k not in old.files or not np.array_equal(np.asarray(g[k]), old[k], equal_nan=True)]
            report.append((os.path.basename(path), len(g), bad))
            print("%-18s %3d arrays: %s" % (os.path.basename(path), len(g), "identical" if not bad else "DIFFER: " + ", ".join(bad)),
                  flush=True)
        np.savez_compressed = compare
    else:
        ref = refload.load()
    src, tran, surf, anal = ref.sources, ref.transformations, ref.surfaces, ref.analyses
    out = {}

    # ---- config 1: Wolter-I pair, on axis (singlePassAlignment.py:246-269)
    N = 4000
    np.random.seed(0)
    st = np.random.get_state()
    u = np.random.rand(2 * N)
    np.random.set_state(st)
    rays = src.subannulus(220., 220.6, 2 * np.pi, N, zhat=-1.)
    g = dict(u1=u[:N], u2=u[N:])
    g.update(pack("rays_in", rays))
    tran.transform(rays, 0, 0, -8400., 0, 0, 0)
    surf.wolterprimary(rays, 220., 8400.)
    g.update(pack("after_primary", rays))
    tran.reflect(rays)
    surf.woltersecondary(rays, 220., 8400.)
    tran.reflect(rays)
    surf.flat(rays)
    g.update(pack("rays_out", rays))
    g["hpd"] = anal.hpd(rays)
    g["rms"] = anal.rmsCentroid(rays)
    g["centroid"] = np.array(anal.centroid(rays))
    w = np.linspace(.5, 1.5, N)
    g["weights"] = w
    g["hpd_w"] = anal.hpd(rays, weights=w)
    g["rms_w"] = anal.rmsCentroid(rays, weights=w)
    g["centroid_w"] = np.array(anal.centroid(rays, weights=w))
    r, cdf = anal.rhocdf(rays, weights=w)
    g["rhocdf_r"], g["rhocdf_cdf"] = r, cdf
    np.savez_compressed(os.path.join(HERE, "wolter1.npz"), **g)

    # ---- config 2: Wolter-Schwarzschild, 10 arcmin off axis (axialHeights.py:77-113)
    N = 4000
    a0, a1 = chains.ws_aperture()
    theta = 10. / 60. * np.pi / 180.
    np.random.seed(0)
    rays = src.subannulus(a0, a1, 100. / 220., N)
    g = dict(a0=a0, a1=a1, theta=theta)
    g.update(pack("rays_in", rays))
    tran.transform(rays, 0, 0, -1.e4, 0, 0, 0)
    surf.wsPrimary(rays, 220., 1.e4, 1.)
    rays[4] = rays[4] + np.sin(theta)
    rays[6] = -np.sqrt(1. - rays[4] ** 2)
    tran.reflect(rays)
    surf.wsSecondary(rays, 220., 1.e4, 1.)
    tran.reflect(rays)
    g.update(pack("after_secondary", rays))
    g["dz_analytic"] = anal.analyticImagePlane(rays)
    g["focus"] = surf.focusI(rays)
    g.update(pack("rays_out", rays))
    g["hpd"] = anal.hpd(rays)
    g["rms"] = anal.rmsCentroid(rays)
    np.savez_compressed(os.path.join(HERE, "ws_offaxis.npz"), **g)

    # ---- config 2b: far off axis (25 arcmin): rays hit the 26-iteration cap and are restored
    theta = 25. / 60. * np.pi / 180.
    np.random.seed(1)
    rays = src.subannulus(a0, a1, 100. / 220., 2000)
    g = dict(theta=theta)
    g.update(pack("rays_in", rays))
    tran.transform(rays, 0, 0, -1.e4, 0, 0, 0)
    surf.wsPrimary(rays, 220., 1.e4, 1.)
    rays[4] = rays[4] + np.sin(theta)
    rays[6] = -np.sqrt(1. - rays[4] ** 2)
    tran.reflect(rays)
    before = copy(rays)
    surf.wsSecondary(rays, 220., 1.e4, 1.)
    g["restored"] = np.logical_and(before[1] == rays[1], np.logical_and(before[2] == rays[2], before[3] == rays[3]))
    g.update(pack("rays_out", rays))
    np.savez_compressed(os.path.join(HERE, "ws_cap.npz"), **g)

    # ---- config 3: Zernike figure error (singlePassAlignment.py:22-56 style) + Wolter-I + vignette
    N = 2000
    ro, ao = chains.zernike_orders(7)
    coeff = chains.zernike_coeff(36, 0)
    np.random.seed(2)
    rays = src.circularbeam(60., N)
    g = dict(coeff=coeff, rorder=ro, aorder=ao, rad=62.5)
    g.update(pack("rays_in", rays))
    tran.transform(rays, 0, 0, -100., 0, 0, 0)
    surf.zernsurf(rays, coeff, 62.5, rorder=ro, aorder=ao, nr=1.)
    g.update(pack("after_zern", rays))
    tran.reflect(rays)
    tran.transform(rays, 0, 0, 50., 0, 0, 0)
    surf.flat(rays, nr=1.)
    g.update(pack("rays_out", rays))
    # non-OPD variant on a fresh bundle
    np.random.seed(3)
    rays = src.circularbeam(60., N)
    g.update(pack("rays_in2", rays))
    tran.transform(rays, 1., -2., -100., 1e-3, -2e-3, .3)
    surf.zernsurf(rays, coeff, 62.5, rorder=ro, aorder=ao)
    g.update(pack("rays_out2", rays))
    np.savez_compressed(os.path.join(HERE, "zernike.npz"), **g)

    # ---- config 4: SPO cones + radial grating with masks + vignette (arcus/cat.py:203-288 style)
    N = 4000
    R0, F = 737., 12.e3
    np.random.seed(4)
    rays = src.subannulus(R0, R0 + .605, 30. / R0, N, zhat=-1.)
    g = dict(R0=R0, F=F)
    g.update(pack("rays_in", rays))
    tran.transform(rays, 0, 0, 0, 0, 0, .01)
    surf.spoPrimary(rays, R0, F)
    tran.reflect(rays)
    surf.spoSecondary(rays, R0, F)
    tran.reflect(rays)
    g.update(pack("after_spo", rays))
    # move to a grating 200 mm above the focus, hub further down +y
    tran.transform(rays, 0, 0, -(F - 200.), 0, 0, 0)
    tran.transform(rays, R0 * 200. / F * 0 + 0., 0, 0, 0, 0, 0)
    surf.flat(rays)
    tran.transform(rays, 0, 11832.911 - 0., 0, 0, 0, 0)      # origin to the hub
    mask = rays[1] > np.median(rays[1])
    g["mask"] = mask
    tran.reflect(rays, ind=mask)
    tran.radgrat(rays, 160. / 11832.911, -3, 2.4, ind=mask)
    wave = np.random.uniform(3.6, 7.2, N)
    g["wave"] = wave
    tran.radgrat(rays, 160. / 11832.911, 1, wave, ind=~mask)
    g.update(pack("after_grat", rays))
    # evanescent diffraction -> NaN n for part of the bundle (transformationsf.f95:231-234)
    evan = np.arange(N) % 7 == 0
    g["evan"] = evan
    tran.radgrat(rays, 160. / 11832.911, 150, 2.4, ind=evan)
    # a small sphere most rays miss -> direction zeroed (surfacesf.f95:342-345)
    g.update(pack("after_evan", rays))
    g.update(pack("vignetted_evan", tran.vignette(rays)))      # NaN > .1 is False: removed
    tran.transform(rays, np.mean(rays[1]), np.mean(rays[2]), 0, 0, 0, 0)
    surf.conic(rays, 10., 0.)
    g.update(pack("after_miss", rays))
    v = tran.vignette(rays)
    g.update(pack("vignetted", v))
    keep = np.logical_and(rays[1] > 0., np.abs(rays[2]) < 10.)
    g["keep"] = keep
    g.update(pack("vignetted_mask", tran.vignette(rays, ind=keep)))
    np.savez_compressed(os.path.join(HERE, "spo_grating.npz"), **g)

    # ---- remaining routines: conic(+opd), refract, woltersine, wolterprimaryopd, itransform, grat, flat(ind)
    N = 1500
    np.random.seed(5)
    rays = src.pointsource(.02, N)
    g = {}
    g.update(pack("rays_in", rays))
    tran.transform(rays, 0, 0, -500., 0, 0, 0)
    surf.conic(rays, 1000., -1.)
    g.update(pack("after_conic", rays))
    tran.refract(rays, 1., 1.5)
    g.update(pack("after_refract", rays))
    tran.transform(rays, 0, 0, 10., 0, 0, 0)
    surf.conic(rays, -800., .3, nr=1.5)
    g.update(pack("after_conicopd", rays))
    tran.refract(rays, 1.5, 1.)
    tran.itransform(rays, .3, -.2, 5., .01, .02, -.03)
    g.update(pack("after_itransform", rays))
    sel = np.arange(0, N, 3)
    g["sel"] = sel
    surf.flat(rays, ind=sel)
    g.update(pack("after_flat_ind", rays))
    order = np.random.randint(-2, 3, N).astype(np.float64)
    wv = np.random.uniform(1., 2., N)
    g["grat_order"], g["grat_wave"] = order, wv
    order[::50] = 150.       # evanescent: l**2+m**2 > 1 -> direction zeroed (transformationsf.f95:297-301)
    g["grat_order"] = order
    tran.grat(rays, 160.e-6, order, wv * 1.e-6)
    g.update(pack("after_grat", rays))
    np.random.seed(6)
    rays = src.subannulus(220., 220.4, .2, N, zhat=-1.)
    g.update(pack("rays_in2", rays))
    tran.transform(rays, 0, 0, -8400., 0, 0, 0)
    surf.woltersine(rays, 220., 8400., 1.e-4, 1. / 20.)
    g.update(pack("after_woltersine", rays))
    np.random.seed(7)
    rays = src.subannulus(220., 220.4, .2, N, zhat=-1.)
    tran.transform(rays, 0, 0, -8400., 0, 0, 0)
    surf.wolterprimary(rays, 220., 8400., psi=1.3, nr=1.)
    g.update(pack("after_primaryopd", rays))
    np.savez_compressed(os.path.join(HERE, "misc.npz"), **g)
    # ---- Legendre-Legendre shells (examples/axro/singlePassAlignment.py:80,157,169 style)
    N = 2000
    rng = np.random.default_rng(8)
    axial = np.array([0, 1, 2, 3, 1, 2, 4, 0, 5])
    az = np.array([0, 0, 0, 1, 1, 2, 2, 3, 1])
    coeff = rng.normal(0., 2.e-4, axial.size)
    g = dict(coeff=coeff, axial=axial, az=az)
    np.random.seed(8)
    rays = src.subannulus(220., 220.6, .25, N, zhat=-1.)
    g.update(pack("rays_in", rays))
    tran.transform(rays, 0, 0, -8400., 0, 0, 0)
    surf.primaryLL(rays, 220., 8400., 8500., 8400., .25, coeff, axial, az)
    g.update(pack("after_primaryLL", rays))
    tran.reflect(rays)
    surf.secondaryLL(rays, 220., 8400., 1., 8400., 8300., .25, coeff * .5, axial, az)
    g.update(pack("after_secondaryLL", rays))
    tran.reflect(rays)
    surf.flat(rays)
    g.update(pack("rays_out", rays))
    g["hpd"] = anal.hpd(rays)
    # ellipsoid-hyperboloid with L-L terms, finite source distance S
    np.random.seed(9)
    rays = src.subannulus(220., 220.5, .25, N, zhat=-1.)
    S = 2.5e5
    # rays diverge from a point source at distance S above the node plane
    tran.transform(rays, 0, 0, -8400. - 60., 0, 0, 0)
    tran.pointTo(rays, 0., 0., 8400. + S, reverse=1.)
    g.update(pack("rays_in2", rays))
    surf.ellipsoidPrimaryLL(rays, 220., 8400., S, 1., 8500., 8400., .25, coeff, axial, az)
    g.update(pack("after_ellipsoidLL", rays))
    tran.reflect(rays)
    surf.ellipsoidSecondaryLL(rays, 220., 8400., S, 1., 8400., 8300., .25, coeff * .5, axial, az)
    tran.reflect(rays)
    surf.flat(rays)
    g.update(pack("rays_out2", rays))
    g["S"] = S
    # plain ellipsoid pair (conic + Wolter secondary with an effective psi)
    np.random.seed(10)
    rays = src.subannulus(220., 220.5, .25, N, zhat=-1.)
    tran.transform(rays, 0, 0, -8400. - 60., 0, 0, 0)
    tran.pointTo(rays, 0., 0., 8400. + S, reverse=1.)
    g.update(pack("rays_in3", rays))
    surf.ellipsoidPrimary(rays, 220., 8400., S, 1.)
    tran.reflect(rays)
    surf.ellipsoidSecondary(rays, 220., 8400., S, 1.)
    tran.reflect(rays)
    surf.flat(rays)
    g.update(pack("rays_out3", rays))
    g["hpd3"] = anal.hpd(rays)
    np.savez_compressed(os.path.join(HERE, "legendre.npz"), **g)

    # ---- the remaining surfaces.py wrappers, each through the reference's own
    # wrapper: sphere / cyl / cylconic / conicplus / torus / paraxial / legSurf / W-S back surfaces /
    # zernphase / zernsurfrot
    N = 1500
    g = {}

    def beam(seed, rad=20., tilt=.02):
        np.random.seed(seed)
        r = src.circularbeam(rad, N)
        rng = np.random.default_rng(seed)
        r[4][:] = rng.normal(0., tilt, N)
        r[5][:] = rng.normal(0., tilt, N)
        r[6][:] = np.sqrt(1. - r[4] ** 2 - r[5] ** 2)
        r[0][:] = rng.normal(0., 1., N)
        return r

    rays = beam(20)
    tran.transform(rays, 0, 0, 300., 0, 0, 0)
    g.update(pack("sphere_in", rays))
    surf.sphere(rays, 250.)
    g.update(pack("sphere_out", rays))
    rays = beam(21)
    g.update(pack("tansphere_in", rays))
    surf.tanSphere(rays, -500., nr=1.5)
    g.update(pack("tansphere_out", rays))
    rays = beam(22)
    tran.transform(rays, 0, 0, 120., 0, 0, 0)
    g.update(pack("cyl_in", rays))
    surf.cyl(rays, 100., nr=1.2)
    g.update(pack("cyl_out", rays))
    rays = beam(23, rad=5.)
    tran.transform(rays, 0, 0, 0, np.pi / 2 - .3, 0, 0)      # grazing onto the y-sag cylinder
    tran.transform(rays, 0, 20., 0, 0, 0, 0)
    g.update(pack("cylconic_in", rays))
    surf.cylconic(rays, 1. / 400., -.7)
    g.update(pack("cylconic_out", rays))
    rays = beam(24)
    tran.transform(rays, 0, 0, 50., 0, 0, 0)
    g.update(pack("conicplus_in", rays))
    pp = np.array([1.e-5, -2.e-9, 3.e-13])
    g["conicplus_p"] = pp
    surf.conicplus(rays, 800., -1.3, pp, nr=1.1)
    g.update(pack("conicplus_out", rays))
    rays = beam(25, rad=8.)
    tran.transform(rays, 0, 0, 30., 0, 0, 0)
    g.update(pack("torus_in", rays))
    surf.torus(rays, 150., 900.)
    g.update(pack("torus_out", rays))
    rays = beam(26)
    g.update(pack("paraxial_in", rays))
    surf.paraxial(rays, 350.)
    surf.paraxialY(rays, -120.)
    g.update(pack("paraxial_out", rays))
    rays = beam(27)
    xo = np.array([0, 1, 2, 3, 1, 2, 5, 0])
    yo = np.array([1, 0, 1, 2, 4, 2, 0, 6])
    lc = np.random.default_rng(27).normal(0., 1.e-4, xo.size)
    g["leg_coeff"], g["leg_xo"], g["leg_yo"] = lc, xo, yo
    g.update(pack("legsurf_in", rays))
    surf.legSurf(rays, 25., 30., 2., lc, xo, yo)
    g.update(pack("legsurf_out", rays))
    # W-S back surfaces: the aperture of config 2 moved out by the substrate thickness
    a0, a1 = chains.ws_aperture()
    np.random.seed(28)
    rays = src.subannulus(a0 + .4, a1 + .4, .3, N, zhat=-1.)
    tran.transform(rays, 0, 0, -1.e4, 0, 0, 0)
    g.update(pack("wsback_in", rays))
    surf.wsPrimaryB(rays, 220., 1.e4, 1., .4)
    g.update(pack("wsback_primary", rays))
    tran.reflect(rays)
    surf.wsSecondaryB(rays, 220., 1.e4, 1., .4)
    g.update(pack("wsback_secondary", rays))
    # Zernike phase screen and the two-set rotated Zernike surface
    rorder = np.array([0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4])
    aorder = np.array([0, 1, -1, 0, -2, 2, -1, 1, -3, 3, 0, 2, -2, 4, -4])
    zc = np.random.default_rng(29).normal(0., 1.e-4, rorder.size)
    zc2 = np.random.default_rng(30).normal(0., 5.e-5, 10)
    g["z_rorder"], g["z_aorder"], g["z_coeff"], g["z_coeff2"] = rorder, aorder, zc, zc2
    rays = beam(29, rad=18.)
    g.update(pack("zernphase_in", rays))
    surf.zernphase(rays, zc, 20., 5.e-4, rorder=rorder, aorder=aorder)
    g.update(pack("zernphase_out", rays))
    rays = beam(30, rad=18.)
    tran.transform(rays, 0, 0, 10., 0, 0, 0)
    g.update(pack("zernrot_in", rays))
    surf.zernsurfrot(rays, zc, zc2, 20., .37, rorder1=rorder, aorder1=aorder, rorder2=rorder[:10], aorder2=aorder[:10])
    g.update(pack("zernrot_out", rays))
    np.savez_compressed(os.path.join(HERE, "surfaces2.npz"), **g)

    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("  %-20s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
