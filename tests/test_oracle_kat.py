"""Known-answer checks that pin the CPU oracle itself.

The reference ships no tests or golden vectors and its Fortran cannot be compiled here, so the
oracle (a C restatement of the .f95 loops) is pinned by physics: closed forms from the
reference's own conicsolve.py, algebraic identities, and the three independent Zernike
implementations inside specialFunctions.f95.
"""
import numpy as np
import pytest

from oracle import chains, pyref
from oracle import f2py as of

T, S, W, Z, SP = of.transformationsf, of.surfacesf, of.woltsurf, of.zernsurf, of.specialfunctions


def _vs(r0, z0, psi=1.):
    alpha = .25 * np.arctan(r0 / z0)
    thetah = 2 * (1 + 2 * psi) / (1 + psi) * alpha
    thetap = 2 * psi / (1 + psi) * alpha
    p = z0 * np.tan(4 * alpha) * np.tan(thetap)
    d = z0 * np.tan(4 * alpha) * np.tan(4 * alpha - thetah)
    e = np.cos(4 * alpha) * (1 + np.tan(4 * alpha) * np.tan(thetah))
    return p, d, e


def primrad(z, r0, z0, psi=1.):          # conicsolve.py:7-15
    p, d, e = _vs(r0, z0, psi)
    return np.sqrt(p ** 2 + 2 * p * z + (4 * e ** 2 * p * d) / (e ** 2 - 1))


def secrad(z, r0, z0, psi=1.):           # conicsolve.py:29-37
    p, d, e = _vs(r0, z0, psi)
    return np.sqrt(e ** 2 * (d + z) ** 2 - z ** 2)


def test_K1_on_axis_wolter_focus_and_real4_delta():
    """Perfect on-axis Wolter-I focuses to a point; the spot size that remains is the REAL*4
    rounding of flat's propagation distance (surfacesf.f95:4-29): 1.278e-5 mm at 1e5 rays,
    seed 0.  A pure-double flat would give ~8e-11."""
    rays = chains.wolter1_source(100_000, 0)
    h = chains.wolter1_cpu(rays)
    assert h == pytest.approx(1.278e-5, rel=2e-3)
    # the same chain with a double-precision delta
    rays = chains.wolter1_source(100_000, 0)
    chains.run_steps_cpu(rays, chains.wolter1_steps()[:-1])
    delta = -rays[3] / rays[6]
    rays[1] += delta * rays[4]
    rays[2] += delta * rays[5]
    assert pyref.hpd(rays) < 1e-9


def test_K2_rays_land_on_prescription_radii():
    rays = chains.wolter1_source(20_000, 1)
    chains.run_steps_cpu(rays, chains.wolter1_steps()[:2])
    r = np.hypot(rays[1], rays[2])
    assert np.abs(r - primrad(rays[3], 220., 8400.)).max() < 1e-12 * 220
    chains.run_steps_cpu(rays, chains.wolter1_steps()[2:4])
    r = np.hypot(rays[1], rays[2])
    assert np.abs(r - secrad(rays[3], 220., 8400.)).max() < 1e-10 * 220
    for k in (7, 8, 9):
        assert np.isfinite(rays[k]).all()
    assert np.abs(rays[7] ** 2 + rays[8] ** 2 + rays[9] ** 2 - 1).max() < 1e-15


def test_K3_primary_only_focus_distance():
    """After the paraboloid alone the rays cross the axis at primfocus (conicsolve.py:62-64)."""
    rays = chains.wolter1_source(5_000, 2)
    chains.run_steps_cpu(rays, chains.wolter1_steps()[:3])
    alpha, p, d, e = pyref.woltparam(220., 8400.)
    zf = 8400. + 2 * e ** 2 * d / (e ** 2 - 1)       # distance of the primary focus from the node plane
    # axial crossing: t with x + l t = 0
    t = -rays[1] / rays[4]
    zc = rays[3] + rays[6] * t
    assert np.abs(zc - (8400. - zf)).max() < 1e-6


def test_K4_ws_offaxis_rms_vs_chase_vanspeybroeck():
    """W-S off-axis blur is of the size Chase & Van Speybroeck Eq. 13 predicts (conicsolve.py:250-254);
    SURVEY.md 3.2 measured rms/Z0 = 3.01e-6 at 5' vs 2.08e-6 closed form."""
    alpha = pyref.woltparam(220., 1.e4)[0]
    out = {}
    for arcmin in (0., 5., 10.):
        th = arcmin / 60. * np.pi / 180.
        rays = chains.ws_source(50_000, 0)
        chains.run_steps_cpu(rays, chains.ws_steps(th))
        f = pyref.focusI(rays)
        out[arcmin] = (f, pyref.rmsCentroid(rays) / 1.e4, pyref.hpd(rays))
    assert abs(out[0.][0]) < 1e-3 and out[0.][2] < 1e-4            # on axis: sharp focus at the origin
    assert out[5.][0] == pytest.approx(3.11, abs=.05)               # SURVEY probe: dz = +3.1149 mm
    assert out[5.][1] == pytest.approx(3.01e-6, rel=.05)
    assert out[10.][1] == pytest.approx(1.85e-5, rel=.05)
    eq13 = .135 * 2 * (np.tan(5. / 60 * np.pi / 180) ** 2 / np.tan(alpha)) * 200. / 1.e4
    assert .5 < out[5.][1] / eq13 < 3.


def test_K5_conic_against_closed_form():
    """conic lands rays on z = r^2/(R(1+sqrt(1-(1+K)r^2/R^2))) and gives a unit normal."""
    rng = np.random.default_rng(3)
    for R, K in ((1000., -1.), (500., 0.), (-800., .3), (300., -2.5)):
        n = 5000
        np.random.seed(3)
        rays = pyref.circularbeam(40., n)
        rays[4][:] = rng.normal(0, .01, n)
        rays[5][:] = rng.normal(0, .01, n)
        rays[6][:] = np.sqrt(1 - rays[4] ** 2 - rays[5] ** 2)
        rays[3][:] = -20.
        S.conic(*rays[1:], R, K)
        r2 = rays[1] ** 2 + rays[2] ** 2
        sag = r2 / (R * (1 + np.sqrt(1 - (1 + K) * r2 / R ** 2)))
        # K=-1 with near-axial rays: the reference's quadratic has denom = l^2+m^2 ~ 1e-4 and
        # cancels catastrophically (surfacesf.f95:319-326); that loss is part of its answer
        assert np.abs(rays[3] - sag).max() < (1e-5 if K == -1. else 1e-9)
        assert np.abs(rays[7] ** 2 + rays[8] ** 2 + rays[9] ** 2 - 1).max() < 1e-14


def test_K6_itransform_inverts_transform():
    from util import random_bundle
    rays = random_bundle(10_000, 4)
    keep = [r.copy() for r in rays]
    T.transform(*rays[1:], 1., -2., 3., .1, -.2, .3)
    T.itransform(*rays[1:], 1., -2., 3., .1, -.2, .3)
    for k in range(1, 10):
        assert np.abs(rays[k] - keep[k]).max() < 1e-12 * (400. if k < 4 else 1.)


def test_K7_reflect_twice_is_identity_and_norm_preserving():
    from util import random_bundle
    rays = random_bundle(10_000, 5)
    keep = [r.copy() for r in rays]
    T.reflect(*rays[4:])
    assert np.abs(rays[4] ** 2 + rays[5] ** 2 + rays[6] ** 2 - 1).max() < 1e-14
    T.reflect(*rays[4:])
    for k in (4, 5, 6):
        assert np.abs(rays[k] - keep[k]).max() < 1e-14


def test_K8_radgrat_order_zero_and_grating_equation():
    from util import random_bundle
    rng = np.random.default_rng(6)
    n = 5000
    rays = random_bundle(n, 6)
    rays[1][:] = rng.uniform(-40, 40, n)
    rays[2][:] = 11832.911 + rng.uniform(-40, 40, n)
    rays[4][:] = rng.normal(0, .02, n)
    rays[5][:] = rng.normal(0, .02, n)
    rays[6][:] = -np.sqrt(1 - rays[4] ** 2 - rays[5] ** 2)
    keep = [r.copy() for r in rays]
    T.radgrat(rays[1], rays[2], rays[4], rays[5], rays[6], 2.4, 160. / 11832.911, 0.)
    assert np.array_equal(rays[4], keep[4]) and np.array_equal(rays[5], keep[5])
    T.radgrat(rays[1], rays[2], rays[4], rays[5], rays[6], 2.4, 160. / 11832.911, -3.)
    d = 160. / 11832.911 * np.hypot(rays[1], rays[2])
    dl, dm = rays[4] - keep[4], rays[5] - keep[5]
    assert np.allclose(np.hypot(dl, dm), 3 * 2.4 / d, rtol=1e-9)
    # the change is perpendicular to the local groove direction (radial from the hub)
    rad = np.stack([rays[1], rays[2]]) / np.hypot(rays[1], rays[2])
    assert np.abs(dl * rad[0] + dm * rad[1]).max() < 1e-7          # pi is REAL*4 in the reference: 8.7e-8 rad of yaw error
    assert (np.sign(rays[6]) == np.sign(keep[6])).all()
    # radgratW takes the sign of n from y
    rays2 = [r.copy() for r in keep]
    rays2[2][::2] *= -1
    T.radgratw(rays2[1], rays2[2], rays2[4], rays2[5], rays2[6], np.full(n, 2.4), 160. / 11832.911, 1.)
    assert (np.sign(rays2[6]) == np.sign(rays2[2])).all()


def test_K9_hpd_of_a_uniform_disc_two_statistics():
    np.random.seed(7)
    rays = pyref.circularbeam(3., 400_000)
    assert pyref.hpd(rays) == pytest.approx(np.sqrt(2.) * 3., rel=5e-3)
    w = np.ones(400_000)
    assert pyref.hpd(rays, weights=w) == pytest.approx(3. * (np.sqrt(.75) - np.sqrt(.25)), rel=5e-3)


def _noll_orders(nmax):
    return chains.zernike_orders(nmax)


def test_K10_zernset_three_implementations_agree():
    """zernset's q-recursion (specialFunctions.f95:142-232) against the closed-form radial
    polynomial (:17-38) and an independent numpy evaluation of Z and its derivatives."""
    ro, ao = _noll_orders(7)
    from math import factorial as f

    def R(n, m, rho):
        return sum((-1) ** k * f(n - k) / (f(k) * f((n + m) // 2 - k) * f((n - m) // 2 - k)) * rho ** (n - 2 * k)
                   for k in range((n - m) // 2 + 1))

    def dR(n, m, rho):
        return sum((-1) ** k * f(n - k) / (f(k) * f((n + m) // 2 - k) * f((n - m) // 2 - k)) * (n - 2 * k)
                   * rho ** (n - 2 * k - 1) for k in range((n - m) // 2 + 1) if n - 2 * k > 0)

    for rho in (.05, .3, .77, 1.):
        for theta in (.1, 2.5, -1.7):
            po, dr, dt = SP.zernset(rho, theta, ro, ao)
            for i, (n, mm) in enumerate(zip(ro, ao)):
                m = abs(int(mm))
                n = int(n)
                norm = np.sqrt(2 * (n + 1))
                assert SP.radialpoly(rho, n, m) == pytest.approx(R(n, m, rho), rel=1e-12, abs=1e-14)
                if mm < 0:
                    zz, zr, zt = norm * R(n, m, rho) * np.sin(m * theta), norm * dR(n, m, rho) * np.sin(m * theta), \
                        norm * R(n, m, rho) * np.cos(m * theta) * m
                elif mm > 0:
                    zz, zr, zt = norm * R(n, m, rho) * np.cos(m * theta), norm * dR(n, m, rho) * np.cos(m * theta), \
                        -norm * R(n, m, rho) * np.sin(m * theta) * m
                else:
                    s = float(np.float32(np.sqrt(np.float32(.5))))      # sqrt(0.5) is REAL*4 in the reference
                    zz, zr, zt = norm * s * R(n, m, rho), norm * s * dR(n, m, rho), 0.
                assert po[i] == pytest.approx(zz, rel=1e-10, abs=1e-11)
                assert dr[i] == pytest.approx(zr, rel=1e-8, abs=1e-9)
                assert dt[i] == pytest.approx(zt, rel=1e-10, abs=1e-11)
    # rho = 0 special cases (:174-184)
    po, dr, dt = SP.zernset(0., .3, ro, ao)
    assert po[0] == pytest.approx(np.sqrt(2.) * float(np.float32(np.sqrt(np.float32(.5)))))
    assert np.isfinite(po).all() and np.isfinite(dr).all()


def test_K11_legendre():
    from numpy.polynomial import legendre as L
    xs = np.linspace(-1, 1, 41)
    for n in range(0, 9):
        c = np.zeros(n + 1)
        c[n] = 1
        for x in xs:
            assert SP.legendre(x, n) == pytest.approx(L.legval(x, c), abs=1e-12)
            assert SP.legendrep(x, n) == pytest.approx(L.legval(x, L.legder(c)) if n else 0., abs=1e-10)
    assert SP.legendre(1.7, 3) == SP.legendre(1., 3)          # clamp: evaluated at sign(x)
    assert SP.legendrep(1.7, 3) == 0.


def test_quirks_real4_literals():
    """REAL*4 literals the reference promotes to double (SURVEY.md 8a-Q 2,3)."""
    assert float(np.float32(np.arccos(np.float32(-1.)))) == 3.1415927410125732
    assert float(np.float32(1e-8)) == 9.99999993922529e-09
    # flat's delta is single precision: positions move by float32(-z/n)
    x = np.array([1.]); y = np.array([2.]); z = np.array([-8400.123456789])
    l = np.array([.01]); m = np.array([.02]); n = np.array([np.sqrt(1 - 5e-4)])
    ux, uy, uz = np.zeros(1), np.zeros(1), np.zeros(1)
    S.flat(x, y, z, l, m, n, ux, uy, uz)
    d32 = float(np.float32(8400.123456789 / n[0]))
    assert x[0] == 1. + d32 * .01 and y[0] == 2. + d32 * .02 and z[0] == 0. and uz[0] == 1.


def test_ws_iteration_cap_semantics():
    """A ray that cannot converge is restored to its entry point and keeps its old normal
    (woltsurf.f95:454-469,562-580)."""
    alpha = pyref.woltparam(220., 1.e4)[0]
    rays = chains.ws_source(3000, 8)
    chains.run_steps_cpu(rays, chains.ws_steps(25. / 60. * np.pi / 180.)[:4])
    before = [r.copy() for r in rays]
    W.wssecondary(*rays[1:], alpha, 1.e4, 1.)
    restored = (rays[1] == before[1]) & (rays[2] == before[2]) & (rays[3] == before[3])
    assert 0 < restored.sum() < 3000
    for k in (7, 8, 9):
        assert np.array_equal(rays[k][restored], before[k][restored])      # normal untouched
    assert not np.array_equal(rays[7][~restored], before[7][~restored])


def test_f2py_style_errors():
    a = np.zeros(10)
    with pytest.raises(ValueError):
        T.reflect(a.astype(np.float32), a, a, a, a, a)
    with pytest.raises(ValueError):
        T.reflect(np.zeros(20)[::2], a, a, a, a, a)
    with pytest.raises(ValueError):
        T.reflect(np.zeros(5), a, a, a, a, a)


def test_legendre_shells_reduce_to_wolter_surfaces_without_terms():
    """With every coefficient zero wolterprimLL / woltersecLL are the plain Wolter-I surfaces
    (same root; psi=1 for the primary), and the ellipsoid pair images a point source at S."""
    rays = chains.wolter1_source(3000, 9, dphi=.3)
    pyref.transform(rays, 0, 0, -8400., 0, 0, 0)
    a, b = [r.copy() for r in rays], [r.copy() for r in rays]
    z3 = np.zeros(3)
    W.wolterprimary(*a[1:], 220., 8400., 1.)
    W.wolterprimll(*b[1:], 220., 8400., 8500., 8400., .3, z3, [0, 1, 2], [0, 1, 1])
    for k in range(1, 10):
        assert np.abs(a[k] - b[k]).max() < 1e-9
    T.reflect(*a[4:]); T.reflect(*b[4:])
    W.woltersecondary(*a[1:], 220., 8400., 1.3)
    W.woltersecll(*b[1:], 220., 8400., 1.3, 8400., 8300., .3, z3, [0, 1, 2], [0, 1, 1])
    for k in range(1, 10):
        assert np.abs(a[k] - b[k]).max() < 1e-6            # woltersecLL stops at |delt| <= 1e-7 (woltsurf.f95:319)


# ---- the remaining surfaces: each pinned by the surface equation it solves
def _golden2():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "surfaces2.npz"))


def _rows(g, key):
    return [np.array(g[key][i]) for i in range(10)]


def _on_line(a, b, tol):
    """Output position b lies on input ray a: (b - a) x dir = 0."""
    dx, dy, dz = b[1] - a[1], b[2] - a[2], b[3] - a[3]
    cx = dy * a[6] - dz * a[5]
    cy = dz * a[4] - dx * a[6]
    cz = dx * a[5] - dy * a[4]
    assert max(np.abs(cx).max(), np.abs(cy).max(), np.abs(cz).max()) < tol


def test_K12_sphere_cylinder_land_on_the_quadric_with_radial_normals():
    g = _golden2()
    a, b = _rows(g, "sphere_in"), _rows(g, "sphere_out")
    assert np.allclose(np.sqrt(b[1] ** 2 + b[2] ** 2 + b[3] ** 2), 250., rtol=0, atol=1e-11)
    assert np.allclose(b[7] * 250., b[1], atol=1e-11) and np.allclose(b[9] * 250., b[3], atol=1e-11)
    _on_line(a, b, 1e-10)
    assert np.array_equal(a[4], b[4])                      # directions untouched
    a, b = _rows(g, "cyl_in"), _rows(g, "cyl_out")
    assert np.allclose(np.sqrt(b[1] ** 2 + b[3] ** 2), 100., rtol=0, atol=1e-11)
    assert np.all(b[8] == 0.)
    _on_line(a, b, 1e-10)
    # OPD grew by nr * path length
    path = np.sqrt((b[1] - a[1]) ** 2 + (b[2] - a[2]) ** 2 + (b[3] - a[3]) ** 2)
    assert np.allclose(np.abs(b[0] - a[0]), 1.2 * path, atol=1e-10)
    # a ray that misses is zeroed and gets a NaN normal (surfacesf.f95:74-80,96-99)
    miss = [np.array([v]) for v in (0., 500., 0., 300., 0., 0., -1., 0., 0., 0.)]
    S.tracesphere(*miss[1:], 250.)
    assert miss[1][0] == 0. and miss[6][0] == 0. and np.isnan(miss[7][0])


def test_K13_tansphere_touches_the_xy_plane_at_the_origin():
    g = _golden2()
    b = _rows(g, "tansphere_out")
    rad = -500.
    # in the tangent frame the sphere is x^2+y^2+(z-rad)^2 = rad^2
    assert np.allclose(b[1] ** 2 + b[2] ** 2 + (b[3] - rad) ** 2, rad ** 2, rtol=1e-13)
    assert np.abs(b[3]).max() < 20. ** 2 / abs(rad)        # sag of a 20 mm beam


def test_K14_cylconic_conicplus_torus_satisfy_their_surface_functions():
    g = _golden2()
    a, b = _rows(g, "cylconic_in"), _rows(g, "cylconic_out")
    c, k = 1. / 400., -.7
    sag = c * b[1] ** 2 / (1 + np.sqrt(1 - (1 + k) * c ** 2 * b[1] ** 2))
    assert np.abs(b[2] - sag).max() < 1e-9                 # Newton tolerance 1e-10 on the step
    assert np.all(b[9] == 0.)
    _on_line(a, b, 1e-9)
    a, b = _rows(g, "conicplus_in"), _rows(g, "conicplus_out")
    R, K, p = 800., -1.3, g["conicplus_p"]
    r = np.sqrt(b[1] ** 2 + b[2] ** 2)
    cc = 1 / R
    surf = cc * r ** 2 / (1 + np.sqrt(1 - (K + 1) * cc ** 2 * r ** 2)) - sum(p[j] * r ** (2 * j + 2) for j in range(3))
    assert np.abs(b[3] - surf).max() < 1e-9
    assert np.allclose(b[7] ** 2 + b[8] ** 2 + b[9] ** 2, 1., atol=1e-14)
    _on_line(a, b, 1e-9)
    a, b = _rows(g, "torus_in"), _rows(g, "torus_out")
    rin, rout = 150., 900.
    t = b[3] + rin + rout
    F = (t ** 2 + b[2] ** 2 + b[1] ** 2 + rout ** 2 - rin ** 2) ** 2 - 4 * rout ** 2 * (b[2] ** 2 + t ** 2)
    assert np.abs(F).max() / (4 * rout ** 2 * rin ** 2) < 1e-9
    # the reference leaves Fz out of the normal's length (surfacesf.f95:499): ux^2+uy^2 = 1 exactly-ish
    assert np.allclose(b[7] ** 2 + b[8] ** 2, 1., atol=1e-14)
    _on_line(a, b, 1e-9)


def test_K15_paraxial_lens_brings_a_collimated_beam_to_its_focus():
    np.random.seed(3)
    rays = pyref.circularbeam(10., 2000)
    S.paraxial(*rays[1:], 350.)
    assert np.allclose(rays[4], -rays[1] / 350.) and np.allclose(rays[5], -rays[2] / 350.)
    assert np.all(rays[6] == 1.)                           # n is left alone (surfacesf.f95:433-437)
    # x + l * 350 = 0: every ray crosses the axis at F (in the reference's small-angle convention)
    assert np.abs(rays[1] + rays[4] * 350.).max() < 1e-12
    r2 = pyref.circularbeam(10., 2000)
    S.paraxialy(*r2[1:], -120.)
    assert np.all(r2[4] == 0.) and np.allclose(r2[5], r2[2] / 120.)


def test_K16_legsurf_against_numpy_legendre():
    from numpy.polynomial import legendre as L
    g = _golden2()
    a, b = _rows(g, "legsurf_in"), _rows(g, "legsurf_out")
    coeff, xo, yo = g["leg_coeff"], g["leg_xo"], g["leg_yo"]
    xw, yw, order = 25., 30., 2.
    dpx = np.zeros_like(a[1]); dpy = np.zeros_like(a[1])
    for c, i, j in zip(coeff, xo, yo):
        ex = np.zeros(i + 1); ex[i] = 1.
        ey = np.zeros(j + 1); ey[j] = 1.
        dpx += c * L.legval(a[2] / yw, ey) * L.legval(a[1] / xw, L.legder(ex))
        dpy += c * L.legval(a[2] / yw, L.legder(ey)) * L.legval(a[1] / xw, ex)
    assert np.allclose(b[4], a[4] + dpx * order / xw, atol=1e-15)
    assert np.allclose(b[5], a[5] + dpy * order / yw, atol=1e-15)
    assert np.allclose(b[4] ** 2 + b[5] ** 2 + b[6] ** 2, 1., atol=1e-15)


def test_K17_ws_back_surface_is_the_front_surface_moved_out_by_the_thickness():
    """A ray at radius r on the back surface solves the front-surface equation at radius r-thick
    (woltsurf.f95:749-753): tracing the same axial rays to the front surface from an aperture
    shifted inwards by `thick` must give the same z."""
    a0, a1 = chains.ws_aperture()
    alpha = pyref.woltparam(220., 1.e4)[0]
    n = 400
    rad = np.linspace(a0 + .05, a1 - .05, n)
    def axial(r):
        z = np.zeros(n)
        return [z.copy(), r.copy(), z.copy(), np.full(n, 1.e4 + 300.), z.copy(), z.copy(),
                np.full(n, -1.), z.copy(), z.copy(), z.copy()]
    front = axial(rad)
    back = axial(rad + .4)
    W.wsprimary(*front[1:], alpha, 1.e4, 1.)
    W.wsprimaryback(*back[1:], alpha, 1.e4, 1., .4)
    assert np.abs(front[3] - back[3]).max() < 1e-7
    assert np.allclose(back[1], rad + .4)                  # axial rays keep their radius
    assert np.allclose(front[7], back[7], atol=1e-9) and np.allclose(front[9], back[9], atol=1e-9)


def test_K18_zernphase_and_rotated_set():
    g = _golden2()
    ro, ao, zc, zc2 = g["z_rorder"], g["z_aorder"], g["z_coeff"], g["z_coeff2"]
    a, b = _rows(g, "zernphase_in"), _rows(g, "zernphase_out")
    # opd picks up wave * sum(c Z); check against zernset directly
    rho = np.sqrt(a[1] ** 2 + a[2] ** 2) / 20.
    th = np.arctan2(a[2], a[1])
    S0 = np.array([np.dot(zc, SP.zernset(r, t, ro, ao)[0]) for r, t in zip(rho[:200], th[:200])])
    assert np.allclose(b[0][:200] - a[0][:200], 5.e-4 * S0, atol=1e-18)
    assert np.allclose(b[4] ** 2 + b[5] ** 2 + b[6] ** 2, 1., atol=1e-15)
    assert np.array_equal(a[1], b[1])                      # positions untouched
    # rotated second set: rot = 0 must equal the single-set tracezern of the summed coefficients
    r1 = _rows(g, "zernrot_in"); r2 = _rows(g, "zernrot_in")
    c2 = np.zeros_like(zc); c2[:10] = zc2
    Z.tracezernrot(*r1[1:], zc, ro, ao, zc2, ro[:10], ao[:10], 20., 0.)
    Z.tracezern(*r2[1:], zc + c2, ro, ao, 20.)
    for k in range(1, 10):
        assert np.allclose(r1[k], r2[k], atol=1e-12), k
    # and a rotation by 2*pi/m leaves an m-fold term unchanged: rot = pi with only even-m terms
    even = np.array([i for i in range(10) if ao[i] % 2 == 0])
    r3 = _rows(g, "zernrot_in"); r4 = _rows(g, "zernrot_in")
    Z.tracezernrot(*r3[1:], zc, ro, ao, zc2[even], ro[even], ao[even], 20., np.pi)
    Z.tracezernrot(*r4[1:], zc, ro, ao, zc2[even], ro[even], ao[even], 20., 0.)
    for k in range(1, 10):
        assert np.allclose(r3[k], r4[k], atol=1e-12), k


# ---------------------------------------------------------------- reconstruct.f95
def _pad100(a):
    t = np.zeros((a.shape[0] + 2, a.shape[1] + 2), order="F") + 100.
    t[1:-1, 1:-1] = a
    return t


def test_K19_southwell_reconstruction_recovers_a_known_surface():
    """reconstruct (reconstruct.f95:1-128) integrates the gradients of a Legendre surface inside a circular
    aperture back to that surface (up to piston and the reference's sign convention, southwell.py:45-46)."""
    n = 48
    xg, yg = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-1, 1, n))
    img = np.polynomial.legendre.legval2d(xg, yg, [[0, 1, 0], [0, .5, 0], [1, 0, 0]])
    gx, gy = np.gradient(img)
    out = np.sqrt(xg ** 2 + yg ** 2) > 1
    gx[out] = 100.
    gy[out] = 100.
    phase = np.zeros(gx.shape, order="F")
    phase[out] = 100.
    P, GX, GY = _pad100(phase), _pad100(gx), _pad100(gy)
    pc = of.reconstruct.reconstruct(GX, GY, 1e-10, 1., P, 10000)
    assert 50 < of.reconstruct.reconstruct.sweeps < 10000          # converged, not capped
    assert np.array_equal(pc, P)                                   # phase is updated in place (:114)
    res = -pc[1:-1, 1:-1]
    assert (res[out] == -100.).all()
    d = (res - img)[~out]
    assert np.std(d) < 1e-7 * np.std(img[~out])


def test_K20_southwellbin_mean_slopes_and_empty_lenslets():
    """southwellbin (reconstruct.f95:136-187): every lenslet holds tan(asin(mean direction cosine)) of its rays,
    empty lenslets are flagged 100., for even and odd array sizes (the two index formulas, :155-164)."""
    rng = np.random.default_rng(5)
    n = 5000
    x, y = rng.uniform(-9.9, 9.9, n), rng.uniform(-4.9, 4.9, n)
    l, m = 1e-3 * x, -2e-3 * y
    for xd, yd, bs in ((20, 10, 1.), (21, 11, 1.)):
        xa, ya, ph = of.reconstruct.southwellbin(x, y, l, m, bs, xd, yd)
        if xd % 2 == 0:
            xb, yb = np.floor(x / bs).astype(int) + xd // 2 + 1, np.floor(y / bs).astype(int) + yd // 2 + 1
        else:
            xb, yb = np.floor((x + bs / 2) / bs).astype(int) + (xd - 1) // 2, np.floor((y + bs / 2) / bs).astype(int) + (yd - 1) // 2
        for cx, cy in ((3, 4), (xd // 2, yd // 2), (xd - 3, 2)):
            sel = (xb == cx) & (yb == cy)
            if sel.any():
                assert xa[cx, cy] == pytest.approx(np.tan(np.arcsin(l[sel].mean())), rel=1e-12)
                assert ya[cx, cy] == pytest.approx(np.tan(np.arcsin(m[sel].mean())), rel=1e-12)
                assert ph[cx, cy] == 0.
        assert (xa[0, :] == 100.).all() or xd % 2 == 1             # even sizes never fill row 0 (:156: +1+1)
        empty = ph == 100.
        assert np.array_equal(empty, xa == 100.) and np.array_equal(empty, ya == 100.)
