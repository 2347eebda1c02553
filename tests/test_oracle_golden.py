"""The oracle restatement (oracle/pyref.py + oracle/chains.py over the C oracle) against
(a) the committed golden vectors, which were produced by the reference's own unmodified
Python layer (tests/golden/make_golden.py), and (b) the live reference Python when
/root/reference is present (build container only)."""
import numpy as np
import pytest

from oracle import chains, pyref, refload
from oracle import f2py as of
from util import assert_bit_equal, copy, rows_of


def test_golden_wolter1_reproduced_by_restatement(golden):
    g = golden("wolter1")
    np.random.seed(0)
    rays = pyref.subannulus(220., 220.6, 2 * np.pi, 4000, zhat=-1.)
    assert_bit_equal(rays, rows_of(g["rays_in"]), what="subannulus")
    assert np.array_equal(g["u1"], np.random.RandomState(0).rand(8000)[:4000])
    chains.run_steps_cpu(rays, chains.wolter1_steps()[:2])
    assert_bit_equal(rays, rows_of(g["after_primary"]), what="after_primary")
    chains.run_steps_cpu(rays, chains.wolter1_steps()[2:])
    assert_bit_equal(rays, rows_of(g["rays_out"]), what="rays_out")
    w = g["weights"]
    assert pyref.hpd(rays) == g["hpd"] and pyref.rmsCentroid(rays) == g["rms"]
    assert pyref.hpd(rays, weights=w) == g["hpd_w"] and pyref.rmsCentroid(rays, weights=w) == g["rms_w"]
    assert np.array_equal(np.array(pyref.centroid(rays, weights=w)), g["centroid_w"])
    r, cdf = pyref.rhocdf(rays, weights=w)
    assert np.array_equal(r, g["rhocdf_r"]) and np.array_equal(cdf, g["rhocdf_cdf"])


def test_golden_ws_reproduced_by_restatement(golden):
    g = golden("ws_offaxis")
    a0, a1 = chains.ws_aperture()
    assert a0 == g["a0"] and a1 == g["a1"]
    assert a0 == 220.13737656836065 and a1 == 221.23348169132342       # SURVEY.md 3.2 probe
    rays = chains.ws_source(4000, 0)
    assert_bit_equal(rays, rows_of(g["rays_in"]), what="ws rays_in")
    chains.run_steps_cpu(rays, chains.ws_steps(float(g["theta"])))
    assert_bit_equal(rays, rows_of(g["after_secondary"]), what="ws after_secondary")
    assert pyref.analyticImagePlane(rays) == g["dz_analytic"]
    assert pyref.focusI(rays) == g["focus"]
    assert_bit_equal(rays, rows_of(g["rays_out"]), what="ws rays_out")
    assert pyref.hpd(rays) == g["hpd"] and pyref.rmsCentroid(rays) == g["rms"]


def test_golden_vignette_reproduced_by_restatement(golden):
    g = golden("spo_grating")
    assert_bit_equal(pyref.vignette(rows_of(g["after_evan"])), rows_of(g["vignetted_evan"]))
    assert_bit_equal(pyref.vignette(rows_of(g["after_miss"])), rows_of(g["vignetted"]))
    assert_bit_equal(pyref.vignette(rows_of(g["after_miss"]), ind=g["keep"]), rows_of(g["vignetted_mask"]))
    # masked gather/scatter idiom
    rays = rows_of(g["after_grat"])
    pyref.masked(of.transformationsf.radgrat, [rays[1], rays[2], rays[4], rays[5], rays[6]], g["evan"],
                 2.4, 160. / 11832.911, 150)
    assert_bit_equal(rays, rows_of(g["after_evan"]), what="masked radgrat")


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_restatement_matches_live_reference_python():
    """Every numpy restatement in oracle/pyref.py against the reference's own function."""
    ref = refload.load()
    src, tran, anal, surf, con = ref.sources, ref.transformations, ref.analyses, ref.surfaces, ref.conicsolve
    for name, args in (("subannulus", (220., 220.6, 1.3, 5001, -1.)), ("annulus", (200., 230., 5001, 1.)),
                       ("circularbeam", (12.5, 5001)), ("pointsource", (.03, 5001))):
        np.random.seed(11)
        a = getattr(src, name)(*args)
        np.random.seed(11)
        b = getattr(pyref, name)(*args)
        assert_bit_equal(b, a, what=name)
    # the set-up sources (the product generates them on the device; tests/test_gpu_parity.py compares with these)
    for name, args in (("xslit", (-3., 5., 257, 1.)), ("rectArray", (4., 2.5, 33)),
                       ("convergingbeam", (8400., 200., 230., -.1, .3, 4001, 1.5)),
                       ("convergingbeam2", (8400., -20., 30., 190., 240., 4001, .5)),
                       ("rectbeam", (12., 7., 4001)), ("gaussianBeam", (.01, 4001)),
                       ("fanBeam", (.02, .03, 21)), ("circFan", (.05, 7, 12))):
        np.random.seed(13)
        a = getattr(src, name)(*args)
        np.random.seed(13)
        b = getattr(pyref, name)(*args)
        assert_bit_equal(b, a, what=name)
    rays = chains.wolter1_source(5001, 12)
    chains.run_steps_cpu(rays, chains.wolter1_steps())
    rays[6][::13] = np.nan
    rays[4][::17] = 0.; rays[5][::17] = 0.; rays[6][::17] = 0.
    assert_bit_equal(pyref.vignette(copy(rays)), tran.vignette(copy(rays)), what="vignette")
    good = pyref.vignette(copy(rays))
    w = np.random.default_rng(12).uniform(.2, 3., good[1].size)
    for fn in ("centroid", "rmsCentroid", "hpd", "analyticImagePlane"):
        for ww in (None, w):
            assert np.array_equal(np.array(getattr(pyref, fn)(good, weights=ww)),
                                  np.array(getattr(anal, fn)(good, weights=ww))), fn
    ra, ca = anal.rhocdf(good, weights=w)
    rb, cb = pyref.rhocdf(good, weights=w)
    assert np.array_equal(ra, rb) and np.array_equal(ca, cb)
    assert pyref.woltparam(220., 8400.) == con.woltparam(220., 8400.)
    a, b = copy(good), copy(good)
    assert pyref.focusI(a) == surf.focusI(b)
    assert_bit_equal(a, b, what="focusI")
    # the two scalar set-up formulas of analyses.py
    from pyxfocus_b200 import analyses as panal
    xs, ys = np.linspace(-3., 4., 11), np.linspace(1., 9., 11)
    for u, v in zip(panal.radialGrad(xs, ys, 2e-4, .03, 11832.), anal.radialGrad(xs, ys, 2e-4, .03, 11832.)):
        assert np.array_equal(u, v)
    assert np.array_equal(panal.sellmeier([.5, .6, .7], [1.03, .23, 1.01], [6e-3, 2e-2, 103.]),
                          anal.sellmeier([.5, .6, .7], [1.03, .23, 1.01], [6e-3, 2e-2, 103.]))
    # the mirrored host-side helpers of the product package
    import importlib
    import sys
    import types
    # import the pure-numpy product modules without touching CUDA: they only need _lib lazily
    from pyxfocus_b200 import conicsolve as pcon
    from pyxfocus_b200 import transformations as ptran
    for f, a in (("primrad", (8450., 220., 8400.)), ("secrad", (8350., 220., 8400.)), ("woltparam", (220., 8400.)),
                 ("primfocus", (220., 8400.)), ("wsRMS", (1., 1e-3, 6e-3, 200., 1e4)), ("wsFoc", (3., 1., 200., 1e4, 6e-3))):
        assert np.array_equal(np.array(getattr(pcon, f)(*a)), np.array(getattr(con, f)(*a))), f
    for f, a in (("primsag", (8500., 220., 8400.)), ("secsag", (8300., 8400., 220., 8400.)), ("rGoal_to_rMax", (221., 8400., 8500.)),
                 ("ellipsoidRad", (1e5, 1., 220., 8400., 8450.)), ("ehSecRad", (1e5, 1., 220., 8400., 8350.)),
                 ("ellipsoidSag", (1e5, 1., 220., 8400., 8500., 8400.)),
                 ("solveS", (8500., 9000., 30., .99, 8900., 0., 0., -1., 0., 0., 1.))):
        got, want = np.array(getattr(pcon, f)(*a)), np.array(getattr(con, f)(*a))
        assert np.all(np.isfinite(want)) and np.array_equal(got, want), f
    for inv in (False, True):
        assert np.allclose(ptran.rotationM(.1, -.2, .3, inverse=inv), tran.rotationM(.1, -.2, .3, inverse=inv),
                           rtol=0, atol=1e-16)
    assert np.array_equal(ptran.translationM(1., 2., 3.), tran.translationM(1., 2., 3.))
    c1, c2 = ptran.newCoords(), tran.newCoords()
    ptran._update_coords_fwd(c1, 1., 2., 3., .1, .2, .3)
    rotm, tranm = tran.rotationM(.1, .2, .3), tran.translationM(1., 2., 3.)
    assert np.allclose(c1[1], np.dot(np.dot(rotm, tranm), c2[1]), atol=1e-15)


def test_golden_southwell_reproduced_by_the_oracle(golden):
    """tests/golden/southwell.npz was written by the reference's unmodified southwell.py driving the oracle's
    ``reconstruct``; re-running the oracle through this repository's restatement of the same few lines must give
    the same bits (guards the fixture and the f2py-shaped binding)."""
    g = golden("southwell")
    for tag in ("ex", "ir"):
        gx, gy = g[tag + "_gx"].copy(), g[tag + "_gy"].copy()
        ind = np.isnan(gx) | np.isnan(gy)
        gx[ind] = 100.
        gy[ind] = 100.
        phase = np.zeros(gx.shape, order="F")
        phase[ind] = 100.

        def pad(a):
            t = np.zeros((a.shape[0] + 2, a.shape[1] + 2), order="F") + 100.
            t[1:-1, 1:-1] = a
            return t
        P, GX, GY = pad(phase), pad(gx), pad(gy)
        of.reconstruct.reconstruct(GX, GY, 1e-10, 1., P, 10000 if tag == "ex" else 300)
        assert of.reconstruct.reconstruct.sweeps == int(g[tag + "_sweeps"])
        res = P[1:-1, 1:-1]
        res[ind] = np.nan
        assert np.array_equal(-res, g[tag + "_phase"], equal_nan=True)
    for tag in ("even", "odd"):
        xd, yd, bs = g["bin_%s_dims" % tag]
        xa, ya, ph = of.reconstruct.southwellbin(g["bin_x"], g["bin_y"], g["bin_l"], g["bin_m"], bs, int(xd), int(yd))
        assert np.array_equal(xa, g["bin_%s_xang" % tag]) and np.array_equal(ph, g["bin_%s_phase" % tag])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_interpolatevec_and_wavefront_restatements_match_the_live_reference():
    """oracle/refapi.py interpolateVec against the reference's own function (analyses.py:189-230; both call scipy's
    griddata) bit for bit, and wavefront (:305-334) against the reference's text run with the two things it lacks as
    shipped supplied from outside: ``man.padRect`` (un-vendored; a one-pixel NaN frame) and a ``reconstruct`` whose
    ``maxiter`` has a default (the reference omits the required argument, :327)."""
    pytest.importorskip("scipy.interpolate")
    from oracle import f2py as of, refapi
    ref = refload.load()
    anal = ref.analyses
    rng = np.random.default_rng(4)
    n = 3000
    r, t = 12.5 * np.sqrt(rng.uniform(0, 1, n)), rng.uniform(0, 2 * np.pi, n)
    x, y = r * np.cos(t), r * np.sin(t)
    l, m = 1e-3 * np.sin(x / 5.) + 1e-5 * rng.normal(size=n), 1e-3 * np.cos(y / 7.) + 1e-5 * rng.normal(size=n)
    zero = np.zeros(n)
    rays = [zero.copy(), x, y, zero.copy(), l, m, np.sqrt(1 - l ** 2 - m ** 2), zero.copy(), zero.copy(), zero + 1.]
    for kw in (dict(method="linear"), dict(method="nearest"), dict(method="cubic"), dict(method="linear", polar=True),
               dict(method="linear", xr=[-3., 9.], yr=[-14., 2.], interpVec=x * y)):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = anal.interpolateVec(copy(rays), 4, 31, 23, **kw)
        b = refapi.interpolateVec(copy(rays), 4, 31, 23, **kw)
        assert np.array_equal(a[0], b[0], equal_nan=True) and a[1] == b[1] and a[2] == b[2], kw

    def pad_rect(img):
        out = np.full((img.shape[0] + 2, img.shape[1] + 2), np.nan)
        out[1:-1, 1:-1] = img
        return out
    anal.man.padRect = pad_rect
    anal.reconstruct.reconstruct = lambda xa, ya, crit, h, ph, maxiter=500: of.reconstruct.reconstruct(xa, ya, crit, h, ph, maxiter)
    for method in ("linear", "cubic"):
        a = anal.wavefront(copy(rays), 20, 16, method=method)
        b = refapi.wavefront(copy(rays), 20, 16, method=method, maxiter=500)
        for u, v in zip(a, b):
            assert np.array_equal(u, v, equal_nan=True), method
        assert np.isfinite(a[0]).sum() > 50
