"""GPU parity tests: every CUDA routine (through the C ABI, via the f2py-shaped modules)
against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star / SURVEY.md 7 "FMA"):
  * routines whose per-ray arithmetic is + - * / sqrt only are BIT-EXACT against the oracle
    (libpxf is built -fmad=false, the oracle -ffp-contract=off; scalar-only libm calls are
    made on the host by the same glibc);
  * routines that call libm per ray (atan2/sin/cos/asin/acos/pow) agree to 1e-12: positions
    relative to the system length scale, unit-vector components absolutely;
  * surviving-ray index sets are bit-exact; HPD within 1e-9 relative.
"""
import numpy as np
import pytest

from util import (ROWS, assert_bit_equal, assert_close, chains, copy, of, pyref, random_bundle, rows_of,
                  run_steps_gpu, steps_to_program, to_dev, to_host)

pytestmark = pytest.mark.gpu

N = 50_001          # odd: exercises the double2 tail


@pytest.fixture(scope="module")
def pxf():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pyxfocus_b200
    return pyxfocus_b200


def wolter_inputs(n=N, seed=0):
    rays = chains.wolter1_source(n, seed=seed, dphi=1.0)
    pyref.transform(rays, 0, 0, -8400., 0, 0, 0)
    return rays


def after_primary(n=N, seed=0):
    rays = wolter_inputs(n, seed)
    of.woltsurf.wolterprimary(*rays[1:], 220., 8400., 1.)
    of.transformationsf.reflect(*rays[4:])
    return rays


# --------------------------------------------------------------------------- per routine
ALGEBRAIC = [
    ("transform", lambda: random_bundle(N, 1), [("transform", (1.5, -2.5, 100., .1, -.2, .3))]),
    ("transform_translate_only", lambda: random_bundle(N, 2), [("transform", (0., 0., 8400., 0., 0., 0.))]),
    ("itransform", lambda: random_bundle(N, 3), [("itransform", (1.5, -2.5, 100., .1, -.2, .3))]),
    ("reflect", lambda: random_bundle(N, 4), [("reflect", ())]),
    ("flat", lambda: random_bundle(N, 5), [("flat", ())]),
    ("flatopd", lambda: random_bundle(N, 6), [("flatopd", (1.5,))]),
    ("wolterprimary", wolter_inputs, [("wolterprimary", (220., 8400., 1.))]),
    ("wolterprimary_psi", wolter_inputs, [("wolterprimary", (220., 8400., 1.7))]),
    ("wolterprimaryopd", wolter_inputs, [("wolterprimaryopd", (220., 8400., 1., 1.3))]),
    ("woltersecondary", after_primary, [("woltersecondary", (220., 8400., 1.))]),
    ("conic_parabola", lambda: conic_inputs(7), [("conic", (1000., -1.))]),
    ("conic_sphere_miss", lambda: conic_inputs(8, spread=900.), [("conic", (500., 0.))]),
    ("conicopd", lambda: conic_inputs(9), [("conicopd", (-800., .3, 1.5))]),
    ("spocone", lambda: spo_inputs(10), [("spocone", (737., .25 * np.arctan((737. + .3025) / 12.e3)))]),
    ("spocone_miss", lambda: random_bundle(N, 11), [("spocone", (5., .3))]),
]


def conic_inputs(seed, spread=30.):
    np.random.seed(seed)
    rays = pyref.circularbeam(spread, N)
    rng = np.random.default_rng(seed)
    rays[4][:] = rng.normal(0, .01, N)
    rays[5][:] = rng.normal(0, .01, N)
    rays[6][:] = np.sqrt(1 - rays[4] ** 2 - rays[5] ** 2)
    rays[6][::97] = 1.0          # exercises the K=-1, |n|==1 special case (surfacesf.f95:316-317)
    rays[4][::97] = 0.
    rays[5][::97] = 0.
    rays[3][:] = -50.
    return rays


def spo_inputs(seed):
    np.random.seed(seed)
    rays = pyref.subannulus(737., 737.605, .04, N, zhat=-1.)
    pyref.transform(rays, 0, 0, -300., 0, 0, .01)
    return rays


@pytest.mark.parametrize("name,make,steps", ALGEBRAIC, ids=[a[0] for a in ALGEBRAIC])
def test_algebraic_routines_bit_exact(pxf, name, make, steps):
    cpu = make()
    dev = to_dev(cpu)
    chains.run_steps_cpu(cpu, steps)
    run_steps_gpu(dev, steps)
    assert_bit_equal(to_host(dev), cpu, what=name)


def assert_same_bits(got, want, what=""):
    """Stricter than assert_bit_equal: the 64-bit patterns agree (sign of zero included); NaN
    payloads are not compared, NaN positions are."""
    for k in range(10):
        a, b = np.asarray(got[k]), np.asarray(want[k])
        na, nb = np.isnan(a), np.isnan(b)
        assert np.array_equal(na, nb), "%s row %s: NaN pattern differs" % (what, ROWS[k])
        ia, ib = a.view(np.uint64)[~na], b.view(np.uint64)[~nb]
        bad = np.flatnonzero(ia != ib)
        assert bad.size == 0, "%s row %s: %d bit patterns differ, first %r vs %r" % (
            what, ROWS[k], bad.size, a[~na][bad[0]], b[~nb][bad[0]])


def special_value_bundle(seed, huge=True):
    """Finite rays salted with -0, +0, NaN, +-Inf, denormals and (optionally) huge values in every row."""
    rays = random_bundle(N, seed)
    rng = np.random.default_rng(seed + 1000)
    specials = [-0., 0., np.nan, np.inf, -np.inf, 5e-324, -5e-324, 2.2e-308, -1e-310]
    if huge:
        specials += [1e300, -1e305, 2. ** 1017, -(2. ** 1016)]
    specials = np.array(specials)
    for r in rays:
        idx = rng.choice(N, N // 8, replace=False)
        r[idx] = rng.choice(specials, idx.size)
    return rays


# Identity rotations (angle +-0) skip their arithmetic only for finite, non-(-0) components; the
# sign of every zero and the NaN poisoning of non-finite rays must still be the reference's.
ZERO_ANGLE = [
    ("translate", (0., 0., -8400., 0., 0., 0.)),
    ("translate_negzero_angles", (1., -2., 3., -0., -0., -0.)),
    ("rot_z_only", (0., 0., 0., 0., 0., .3)),
    ("rot_x_only", (5., 0., 0., -.2, 0., -0.)),
    ("rot_y_only", (0., 0., 0., -0., .7, 0.)),
    ("general", (1., 2., 3., .1, -.2, .3)),
]


@pytest.mark.parametrize("name,args", ZERO_ANGLE, ids=[a[0] for a in ZERO_ANGLE])
@pytest.mark.parametrize("routine", ["transform", "itransform"])
def test_transform_special_values_same_bits(pxf, routine, name, args):
    cpu = special_value_bundle(21)
    dev = to_dev(cpu)
    steps = [(routine, args)]
    chains.run_steps_cpu(cpu, steps)
    run_steps_gpu(dev, steps)
    assert_same_bits(to_host(dev), cpu, what=routine + " " + name)
    # and inside a fused program (liveness drops the dead triplets, same per-ray code)
    cpu2 = special_value_bundle(22)
    dev2 = to_dev(cpu2)
    steps2 = [(routine, args), ("reflect", ()), (routine, args)]
    chains.run_steps_cpu(cpu2, steps2)
    steps_to_program(steps2).run(dev2)
    assert_same_bits(to_host(dev2), cpu2, what="fused " + routine + " " + name)


def test_division_helpers_same_bits_on_special_values(pxf):
    """flat divides -z/n: zero dividends of either sign, zero/denormal/huge divisors, NaN and Inf
    must come out with the operator's bits (the helpers only shortcut operands in a safe window)."""
    cpu = special_value_bundle(23)
    dev = to_dev(cpu)
    steps = [("flat", ())]
    chains.run_steps_cpu(cpu, steps)
    run_steps_gpu(dev, steps)
    assert_same_bits(to_host(dev), cpu, what="flat specials")
    # Newton surfaces on garbage: every ray ends in the same state as the oracle's.  (No huge finite
    # values here: the kernels evaluate Fx*l+Fy*m as -2*(x*l+y*m), which is the same double unless
    # 2*x*l overflows while x*l does not -- beyond 8.9e307 the Inf/NaN patterns may differ.)
    for name, a in (("wolterprimary", (220., 8400., 1.)), ("woltersecondary", (220., 8400., 1.))):
        cpu = special_value_bundle(24, huge=False)
        for k in (1, 2, 3):
            cpu[k][::3] = random_bundle(N, 25)[k][::3]
        dev = to_dev(cpu)
        chains.run_steps_cpu(cpu, [(name, a)])
        run_steps_gpu(dev, [(name, a)])
        assert_same_bits(to_host(dev), cpu, what=name + " specials")


def test_grat_bit_exact(pxf):
    cpu = random_bundle(N, 12)
    rng = np.random.default_rng(12)
    order = rng.integers(-2, 3, N).astype(np.float64)
    order[::41] = 400.          # evanescent -> zeroed direction
    wave = rng.uniform(1e-6, 2e-6, N)
    dev = to_dev(cpu)
    of.transformationsf.grat(cpu[1], cpu[2], cpu[4], cpu[5], cpu[6], 1.6e-4, order, wave)
    pxf.transformations.grat(dev, 1.6e-4, order, wave)
    assert_bit_equal(to_host(dev), cpu, what="grat")


LIBM = [
    ("refract", lambda: refract_inputs(13), [("refract", (1., 1.5))], 1.),
    ("refract_out", lambda: refract_inputs(14), [("refract", (1.5, 1.))], 1.),
    ("radgrat", lambda: grating_inputs(15), [("radgrat", (2.4, 160. / 11832.911, -3.))], 1.2e4),
    ("woltersine", wolter_inputs, [("woltersine", (220., 8400., 1.e-4, .05))], 8.4e3),
    ("wsprimary", lambda: ws_inputs(16), [("wsprimary", (pyref.woltparam(220., 1.e4)[0], 1.e4, 1.))], 1.e4),
    ("ws_pair_offaxis", lambda: ws_inputs(17), chains.ws_steps(8. / 60. * np.pi / 180.)[1:], 1.e4),
]


def refract_inputs(seed):
    rays = random_bundle(N, seed)
    # normals within ~40 deg of the direction so that both refraction senses are real
    rng = np.random.default_rng(seed)
    for k in range(3):
        rays[7 + k][:] = rays[4 + k] + .4 * rng.normal(0, 1, N)
    nrm = np.sqrt(rays[7] ** 2 + rays[8] ** 2 + rays[9] ** 2)
    for k in range(3):
        rays[7 + k] /= nrm
    rays[7][::53] *= -1
    rays[8][::53] *= -1
    rays[9][::53] *= -1      # flipped normals (dot<0 branch)
    rays[7][::101] = rays[4][::101]
    rays[8][::101] = rays[5][::101]
    rays[9][::101] = rays[6][::101]    # dot == 1 (or its neighbour): skip branch
    return rays


def grating_inputs(seed):
    rng = np.random.default_rng(seed)
    rays = random_bundle(N, seed)
    rays[1][:] = rng.uniform(-40, 40, N)
    rays[2][:] = 11832.911 + rng.uniform(-40, 40, N)
    rays[4][:] = rng.normal(0, .02, N)
    rays[5][:] = rng.normal(0, .02, N)
    rays[6][:] = -np.sqrt(1 - rays[4] ** 2 - rays[5] ** 2)
    return rays


def ws_inputs(seed):
    rays = chains.ws_source(N, seed=seed)
    pyref.transform(rays, 0, 0, -1.e4, 0, 0, 0)
    return rays


@pytest.mark.parametrize("name,make,steps,scale", LIBM, ids=[a[0] for a in LIBM])
def test_libm_routines_close(pxf, name, make, steps, scale):
    cpu = make()
    dev = to_dev(cpu)
    chains.run_steps_cpu(cpu, steps)
    run_steps_gpu(dev, steps)
    assert_close(to_host(dev), cpu, pos_scale=scale, tol=1e-12, what=name)


def _ws_chain_diff(pxf, arcmin, seed=18):
    cpu = ws_inputs(seed)
    steps = chains.ws_steps(arcmin / 60. * np.pi / 180.)[1:]
    dev = to_dev(cpu)
    chains.run_steps_cpu(cpu, steps)
    run_steps_gpu(dev, steps)
    got = to_host(dev)
    bad = np.zeros(N, bool)
    biteq = np.ones(N, bool)
    for k in range(1, 10):
        scale = 1.e4 if k < 4 else 1.
        both_nan = np.isnan(got[k]) & np.isnan(cpu[k])
        bad |= ~((np.abs(got[k] - cpu[k]) <= 1e-12 * scale) | both_nan)
        biteq &= (got[k] == cpu[k]) | both_nan
    return bad, biteq


@pytest.mark.parametrize("arcmin", [13., 17., 20., 24., 30.])
def test_ws_exact_mode_is_bit_for_bit_at_any_field_angle(pxf, arcmin):
    """Near and beyond the graze angle (18.9') the Newton iteration on the secondary is chaotic for 10-30 % of the
    rays: a 1e-15 change of a ray's state on the primary changes which root it finds or whether the iteration cap
    restores it (measured: a transcendental-free primary in front of a literal secondary moves 9239 of 50001
    outcomes at 17').  The reference's own result for those rays therefore hangs on the last bit of its libm, which
    is unpinned.  PXF_OPT_WS_LIBM evaluates the reference's literal sequence with CORRECTLY ROUNDED sin / cos / tan /
    asin / atan2 / pow (pxf_crmath.cuh) for every ray; against the oracle built the same way (liboracle_cr.so:
    binary128 libm rounded once) the whole primary -> kick -> reflect -> secondary -> reflect chain must then agree
    BIT FOR BIT, chaotic rays included.  Against this image's glibc (correctly rounded for all but ~5e-3 of the
    calls) it agrees for all but a few rays per 1000."""
    pxf.set_option(pxf.OPT_WS_LIBM, 1)
    try:
        with of.libm("cr"):
            bad, biteq = _ws_chain_diff(pxf, arcmin)
        assert (~biteq).sum() == 0, "%g': %d rays differ in some bit from the correctly rounded oracle" % (arcmin, (~biteq).sum())
        bad_g, biteq_g = _ws_chain_diff(pxf, arcmin)
        print("exact mode at %g' vs glibc oracle: %d rays not bit-equal, %d off by > 1e-12" % (arcmin, (~biteq_g).sum(), bad_g.sum()))
        assert bad_g.mean() <= 5e-3
    finally:
        pxf.set_option(pxf.OPT_WS_LIBM, 0)


@pytest.mark.parametrize("arcmin,allowed", [(0., 0.), (6., 0.), (10., 0.), (24., 2e-4), (30., 2e-4), (20., .03), (17., .2)])
def test_ws_far_off_axis_chaotic_fringe(pxf, arcmin, allowed):
    """DEFAULT options (transcendental-free evaluation).  Inside the field of view every ray agrees with the (glibc)
    oracle to 1e-12: 0 of 50001 differ.  In the chaotic band around the graze angle (~12'-21' for this shell) and on
    the fringe of the restored set beyond it the rays that differ are bounded and reported; exact mode (previous
    test) reproduces those bit for bit, fringe mode (next test) the restored sets beyond the band."""
    bad, _ = _ws_chain_diff(pxf, arcmin)
    print("default mode at %g': %d of %d rays off by > 1e-12" % (arcmin, bad.sum(), N))
    assert bad.mean() <= allowed


@pytest.mark.parametrize("arcmin", [20., 24., 30.])
def test_ws_cap_restores_same_rays(pxf, arcmin):
    """Rays that exhaust the 26-iteration cap are restored in place (woltsurf.f95:562-580): the set of restored
    rays is this routine's "surviving-ray index set".  Exact mode: identical ray for ray at every angle.  Fringe mode
    (PXF_OPT_WS_RETRACE = 12): identical beyond the chaotic band (24', 30').  Default: differs on the fringe only
    (< 1 % of the restored rays at 20', ~1e-4 at 24')."""
    a = pyref.woltparam(220., 1.e4)[0]
    for mode in ("exact", "fringe", "default"):
        cpu = ws_inputs(19)
        steps = chains.ws_steps(arcmin / 60. * np.pi / 180.)[1:4]
        chains.run_steps_cpu(cpu, steps)
        dev = to_dev(cpu)
        before = copy(cpu)
        of.woltsurf.wssecondary(*cpu[1:], a, 1.e4, 1.)
        pxf.set_option(pxf.OPT_WS_LIBM, 1 if mode == "exact" else 0)
        pxf.set_option(pxf.OPT_WS_RETRACE, 12 if mode == "fringe" else 0)
        try:
            pxf.woltsurf.wssecondary(*dev[1:], a, 1.e4, 1.)
        finally:
            pxf.set_option(pxf.OPT_WS_LIBM, 0)
            pxf.set_option(pxf.OPT_WS_RETRACE, 0)
        got = to_host(dev)
        rest_cpu = (before[1] == cpu[1]) & (before[2] == cpu[2]) & (before[3] == cpu[3])
        rest_gpu = (before[1] == got[1]) & (before[2] == got[2]) & (before[3] == got[3])
        assert rest_cpu.sum() > 100, "test needs rays that hit the cap"
        ndiff = int((rest_cpu != rest_gpu).sum())
        print("%g' %s mode: restored rays %d (oracle) / %d (GPU), %d differ" % (arcmin, mode, rest_cpu.sum(), rest_gpu.sum(), ndiff))
        if mode == "exact" or (mode == "fringe" and arcmin >= 24.):
            assert ndiff == 0, "%d of %d restored rays differ (%s mode)" % (ndiff, rest_cpu.sum(), mode)
        else:
            assert ndiff <= 1e-2 * rest_cpu.sum(), "%d of %d restored rays differ (%s mode)" % (ndiff, rest_cpu.sum(), mode)


@pytest.mark.parametrize("libm", [0, 1])
def test_ws_both_evaluations_match_oracle(pxf, libm):
    """Config 2 field points inside the field of view (on axis, 3 and 6 arcmin): the default transcendental-free W-S evaluation and
    the libm one both agree with the oracle to 1e-12 through the whole primary -> kick -> secondary chain."""
    pxf.set_option(pxf.OPT_WS_LIBM, libm)
    try:
        for k, arcmin in enumerate((0., 3., 6.)):
            cpu = ws_inputs(40 + k)
            steps = chains.ws_steps(arcmin / 60. * np.pi / 180.)[1:]
            dev = to_dev(cpu)
            chains.run_steps_cpu(cpu, steps)
            run_steps_gpu(dev, steps)
            got = to_host(dev)
            bad = np.zeros(N, bool)
            for r in range(1, 10):
                scale = 1.e4 if r < 4 else 1.
                d = np.abs(got[r] - cpu[r])
                bad |= ~((d <= 1e-12 * scale) | (np.isnan(got[r]) & np.isnan(cpu[r])))
            assert bad.sum() == 0, "%g arcmin: %d rays differ" % (arcmin, bad.sum())
    finally:
        pxf.set_option(pxf.OPT_WS_LIBM, 0)


def test_radgratw_sign_from_y(pxf):
    cpu = grating_inputs(20)
    cpu[2][::3] *= -1          # radgratW takes the sign of n from y (transformationsf.f95:258)
    wave = np.random.default_rng(20).uniform(3.6, 7.2, N)
    dev = to_dev(cpu)
    of.transformationsf.radgratw(cpu[1], cpu[2], cpu[4], cpu[5], cpu[6], wave, 160. / 11832.911, 1.)
    pxf.transformations.radgrat(dev, 160. / 11832.911, 1., wave)
    assert_close(to_host(dev), cpu, pos_scale=1.2e4, what="radgratw")


@pytest.mark.parametrize("nr", [None, 1.5])
def test_fused_zernsurf_bit_identical_to_per_routine(pxf, nr):
    """PXF_OP_ZERNSURF inside a fused program (table staged in shared memory) == transform, zernsurf, reflect,
    flat issued one by one: same device code, same bits; also through `with fused(...)` recording."""
    ro, ao = chains.zernike_orders(7)
    coeff = chains.zernike_coeff(36, 3)
    cpu = random_bundle(N, 77)
    cpu[1] *= .4
    cpu[2] *= .4                                       # inside the unit disc of rad = 62.5 mostly
    T, S = pxf.transformations, pxf.surfaces
    a = to_dev(cpu)
    T.transform(a, 1., -2., -100., 1e-3, -2e-3, .3)
    S.zernsurf(a, coeff, 62.5, rorder=ro, aorder=ao, nr=nr)
    T.reflect(a)
    S.flat(a, nr=nr)
    b = to_dev(cpu)
    before = pxf.launch_count()
    with pxf.fused(b):
        T.transform(b, 1., -2., -100., 1e-3, -2e-3, .3)
        S.zernsurf(b, coeff, 62.5, rorder=ro, aorder=ao, nr=nr)
        T.reflect(b)
        S.flat(b, nr=nr)
    assert pxf.launch_count() - before == 1
    assert_bit_equal(to_host(b), to_host(a), what="fused zernsurf")
    c = to_dev(cpu)
    prog = (pxf.Program().transform(-1., 2., 100., -1e-3, 2e-3, -.3).zernsurf(coeff, ro, ao, 62.5, nr).reflect())
    prog = prog.flat() if nr is None else prog.flatopd(nr)
    prog.run(c)
    assert_bit_equal(to_host(c), to_host(a), what="Program.zernsurf")
    with pytest.raises(ValueError):
        pxf.Program().zernsurf(coeff, ro, ao, 62.5).zernsurf(coeff, ro, ao, 62.5)
    # radial order 8: not fusable -> the recorder flushes and runs the stand-alone routine, same result as unfused
    ro9, ao9 = chains.zernike_orders(8)
    c9 = chains.zernike_coeff(len(ro9), 4)
    d, e = to_dev(cpu), to_dev(cpu)
    T.transform(d, 0, 0, -100., 0, 0, 0)
    S.zernsurf(d, c9, 62.5, rorder=ro9, aorder=ao9)
    with pxf.fused(e):
        T.transform(e, 0, 0, -100., 0, 0, 0)
        S.zernsurf(e, c9, 62.5, rorder=ro9, aorder=ao9)
    assert_bit_equal(to_host(e), to_host(d), what="zernsurf order 8 under fused()")


@pytest.mark.parametrize("opd", [False, True])
def test_tracezern(pxf, opd):
    ro, ao = chains.zernike_orders(7)
    coeff = chains.zernike_coeff(36, 0)
    np.random.seed(21)
    cpu = pyref.circularbeam(60., N)
    pyref.transform(cpu, 1., -2., -100., 1e-3, -2e-3, .3)
    dev = to_dev(cpu)
    if opd:
        of.zernsurf.tracezernopd(*cpu, coeff, ro, ao, 62.5, 1.)
        pxf.surfaces.zernsurf(dev, coeff, 62.5, rorder=ro, aorder=ao, nr=1.)
    else:
        of.zernsurf.tracezern(*cpu[1:], coeff, ro, ao, 62.5)
        pxf.surfaces.zernsurf(dev, coeff, 62.5, rorder=ro, aorder=ao)
    assert_close(to_host(dev), cpu, pos_scale=100., tol=1e-12, what="tracezern")


def test_tracezern_high_order(pxf):
    ro, ao = chains.zernike_orders(11)
    coeff = chains.zernike_coeff(ro.size, 3, sigma=2e-5)
    np.random.seed(22)
    cpu = pyref.circularbeam(55., 20_000)
    pyref.transform(cpu, 0, 0, -10., 0, 0, 0)
    dev = to_dev(cpu)
    of.zernsurf.tracezern(*cpu[1:], coeff, ro, ao, 62.5)
    pxf.surfaces.zernsurf(dev, coeff, 62.5, rorder=ro, aorder=ao)
    assert_close(to_host(dev), cpu, pos_scale=100., tol=1e-12, what="tracezern n<=11")


# --------------------------------------------------------------------------- masks
def test_masked_execution_matches_gather_scatter(pxf):
    """ind= as an in-kernel predicate == the reference's gather -> Fortran -> scatter."""
    rng = np.random.default_rng(23)
    mask = rng.random(N) < .4
    idx = np.where(mask)
    # transform
    cpu = random_bundle(N, 23)
    dev = to_dev(cpu)
    pyref.masked(of.transformationsf.transform, cpu[1:], mask, -1., 2., -3., -.1, .2, -.3)
    pxf.transformations.transform(dev, 1., -2., 3., .1, -.2, .3, ind=mask)
    assert_bit_equal(to_host(dev), cpu, what="transform ind=bool")
    # reflect with an np.where tuple
    pyref.masked(of.transformationsf.reflect, cpu[4:], idx)
    pxf.transformations.reflect(dev, ind=idx)
    assert_bit_equal(to_host(dev), cpu, what="reflect ind=where")
    # flat with an integer index array
    sel = np.arange(0, N, 3)
    pyref.masked(of.surfacesf.flat, cpu[1:], sel)
    pxf.surfaces.flat(dev, ind=sel)
    assert_bit_equal(to_host(dev), cpu, what="flat ind=int")
    # spoCone masked
    cpu = spo_inputs(24)
    dev = to_dev(cpu)
    tg = .25 * np.arctan((737. + .3025) / 12.e3)
    pyref.masked(of.woltsurf.spocone, cpu[1:], mask, 737., tg)
    pxf.surfaces.spoCone(dev, 737., tg, ind=mask)
    assert_bit_equal(to_host(dev), cpu, what="spoCone ind")


def test_unaligned_rows_take_scalar_path(pxf):
    """Rows that are not 16-byte aligned (odd offsets into a larger tensor) must give the
    same bits as the double2 path."""
    import torch
    cpu = wolter_inputs(10_001, 25)
    big = torch.zeros(10, 10_004, dtype=torch.float64, device="cuda")
    dev = [big[i, 1:10_002] for i in range(10)]          # 8-byte offset -> misaligned
    for i in range(10):
        dev[i].copy_(torch.from_numpy(cpu[i]))
    assert dev[1].data_ptr() % 16 == 8
    steps = chains.wolter1_steps()[1:]
    chains.run_steps_cpu(cpu, steps)
    dev_c = [d.contiguous() for d in dev]                # 1-D slices are already contiguous views
    run_steps_gpu(dev_c, steps)
    assert_bit_equal([d.cpu().numpy() for d in dev_c], cpu, what="unaligned")


# --------------------------------------------------------------------------- fused program
def test_fused_program_bit_identical_to_per_routine(pxf):
    cpu = chains.wolter1_source(N, seed=26)
    steps = chains.wolter1_steps()
    a = to_dev(cpu)
    b = to_dev(cpu)
    run_steps_gpu(a, steps)
    steps_to_program(steps).run(b)
    chains.run_steps_cpu(cpu, steps)
    ha, hb = to_host(a), to_host(b)
    assert_bit_equal(hb, ha, what="fused vs per-routine")
    assert_bit_equal(hb, cpu, what="fused vs oracle")


def test_fused_context_records_reference_calls(pxf):
    cpu = chains.wolter1_source(20_000, seed=27)
    dev = to_dev(cpu)
    before = pxf.launch_count()
    with pxf.fused(dev):
        pxf.transformations.transform(dev, 0, 0, -8400., 0, 0, 0)
        pxf.surfaces.wolterprimary(dev, 220., 8400.)
        pxf.transformations.reflect(dev)
        pxf.surfaces.woltersecondary(dev, 220., 8400.)
        pxf.transformations.reflect(dev)
        pxf.surfaces.flat(dev)
    assert pxf.launch_count() - before == 1, "the whole chain must be ONE kernel launch"
    chains.run_steps_cpu(cpu, chains.wolter1_steps())
    assert_bit_equal(to_host(dev), cpu, what="fused context")


def test_fused_all_opcodes(pxf):
    """Every opcode of the fused interpreter against its per-routine kernel (bit-identical:
    same device code) on chains that keep the rays physical."""
    a0 = pyref.woltparam(220., 1.e4)[0]
    cases = {
        "ws": (lambda: ws_inputs(28), chains.ws_steps(6. / 60. * np.pi / 180.)[1:] + [("flatopd", (1.1,))]),
        "conic": (lambda: conic_inputs(29), [("conic", (1000., -1.)), ("refract", (1., 1.5)),
                                            ("transform", (0., 0., -10., 0., 0., 0.)),
                                            ("conicopd", (-800., .3, 1.5)), ("refract", (1.5, 1.)),
                                            ("itransform", (.3, -.2, 5., .01, .02, -.03)), ("flat", ())]),
        "spo": (lambda: spo_inputs(30), [("spocone", (737., .25 * np.arctan((737. + .3025) / 12.e3))), ("reflect", ()),
                                         ("spocone", (737., .75 * np.arctan((737. + .3025) / 12.e3))), ("reflect", ()),
                                         ("transform", (0., -11832.911, 11000., 0., 0., 0.)), ("flat", ()),
                                         ("radgrat", (2.4, 160. / 11832.911, -3.))]),
        "sine": (wolter_inputs, [("woltersine", (220., 8400., 1e-4, .05)), ("reflect", ()),
                                 ("wolterprimaryopd", (220., 8400., 1., 1.))]),
        "wsprim": (lambda: ws_inputs(31), [("wsprimary", (a0, 1.e4, 1.))]),
    }
    for name, (make, steps) in cases.items():
        cpu = make()
        a, b = to_dev(cpu), to_dev(cpu)
        run_steps_gpu(a, steps)
        steps_to_program(steps).run(b)
        assert_bit_equal(to_host(b), to_host(a), what="fused[%s]" % name)


def test_fused_vignette_predicates(pxf):
    """In-program vignetting: rays stop at the predicate; alive flags == the masks the
    reference scripts build (examples/axro/slf.py:145-147)."""
    cpu = chains.wolter1_source(N, seed=32, dphi=1.2)
    dev = to_dev(cpu)
    prog = (pxf.Program().transform(0, 0, 8400., 0, 0, 0).wolterprimary(220., 8400., 1.).reflect()
            .vignette_box(3, 8426., 8526.).vignette_abs(2, 50.)
            .woltersecondary(220., 8400., 1.).reflect().vignette_mag().flat())
    alive = prog.run(dev).cpu().numpy().astype(bool)
    # oracle: same chain with numpy masks
    chains.run_steps_cpu(cpu, chains.wolter1_steps()[:3])
    keep = (cpu[3] > 8426.) & (cpu[3] < 8526.) & (np.abs(cpu[2]) < 50.)
    assert 0 < keep.sum() < N
    assert np.array_equal(alive, keep), "surviving-ray set differs"
    surv = pyref.vignette(cpu, ind=keep)
    chains.run_steps_cpu(surv, chains.wolter1_steps()[3:])
    got = pxf.transformations.vignette(dev, ind=alive)
    assert_bit_equal(to_host(got), surv, what="fused vignette survivors")


# --------------------------------------------------------------------------- vignette / compaction
def test_vignette_default_and_mask(pxf, golden):
    g = golden("spo_grating")
    for src_key, out_key, ind in (("after_evan", "vignetted_evan", None), ("after_miss", "vignetted", None),
                                  ("after_miss", "vignetted_mask", g["keep"])):
        dev = to_dev(rows_of(g[src_key]))
        out = pxf.transformations.vignette(dev, ind=ind)
        assert_bit_equal(to_host(out), rows_of(g[out_key]), what=out_key)


def test_vignette_index_sets_and_order(pxf):
    import torch
    rng = np.random.default_rng(33)
    for n in (1, 31, 32, 33, 2047, 2048, 2049, 300_001):
        cpu = random_bundle(n, 33)
        kill = rng.random(n) < .37
        for k in (4, 5, 6):
            cpu[k][kill] = 0.
        cpu[6][::11] = np.nan
        dev = to_dev(cpu)
        want = pyref.vignette(cpu)
        got = pxf.transformations.vignette(dev)
        assert_bit_equal(to_host(got), want, what="vignette n=%d" % n)
        flags = torch.from_numpy((cpu[4] ** 2 + cpu[5] ** 2 + cpu[6] ** 2 > .1).astype(np.uint8)).cuda()
        idx = pxf.transformations.surviving_indices(flags).cpu().numpy()
        assert np.array_equal(idx, np.where(cpu[4] ** 2 + cpu[5] ** 2 + cpu[6] ** 2 > .1)[0])
    # integer index arrays (np.where output, repeats, negative) behave like numpy fancy indexing
    cpu = random_bundle(1000, 34)
    dev = to_dev(cpu)
    sel = np.array([5, 5, 999, 0, -1, 17])
    assert_bit_equal(to_host(pxf.transformations.vignette(dev, ind=sel)), pyref.vignette(cpu, ind=sel))
    assert pxf.transformations.vignette(dev, ind=np.zeros(1000, bool))[1].shape[0] == 0


# --------------------------------------------------------------------------- analyses
def test_analyses_match_numpy(pxf, golden):
    g = golden("wolter1")
    rays = rows_of(g["rays_out"])
    dev = to_dev(rays)
    A = pxf.analyses
    w = g["weights"]
    cx, cy = A.centroid(dev)
    assert abs(cx - g["centroid"][0]) <= 1e-12 * 220 and abs(cy - g["centroid"][1]) <= 1e-12 * 220
    assert A.rmsCentroid(dev) == pytest.approx(float(g["rms"]), rel=1e-9)
    assert A.hpd(dev) == pytest.approx(float(g["hpd"]), rel=1e-9)
    assert A.hpd(dev, weights=w) == pytest.approx(float(g["hpd_w"]), rel=1e-9)
    assert A.rmsCentroid(dev, weights=w) == pytest.approx(float(g["rms_w"]), rel=1e-9)
    r, cdf = A.rhocdf(dev, weights=w)
    assert np.allclose(r.cpu().numpy(), g["rhocdf_r"], rtol=1e-9, atol=0)
    assert np.allclose(cdf.cpu().numpy(), g["rhocdf_cdf"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 1000, 1001, 65_536, 1_000_003, 2_097_152, 3_000_001, 6_000_000])
def test_hpd_exact_order_statistics(pxf, n):
    """Unweighted hpd == 2*np.median(r): the select must return the exact middle order
    statistics (even n: mean of the two)."""
    rng = np.random.default_rng(n)
    cpu = random_bundle(n, n)
    cpu[1][:] = rng.normal(0, 1e-3, n) + 5.
    cpu[2][:] = rng.standard_cauchy(n) * 1e-3 - 2.
    if n > 100:
        cpu[1][: n // 3] = cpu[1][0]          # many exact ties
        cpu[2][: n // 3] = cpu[2][0]
    dev = to_dev(cpu)
    want = pyref.hpd(cpu)
    got = pxf.analyses.hpd(dev)
    assert got == pytest.approx(want, rel=1e-9)
    # and bit-exact against the median of the radii the device itself computes
    r = pxf.analyses.rho(dev, cent=True).cpu().numpy()
    assert got == 2. * np.median(r)


def test_hpd_bracket_fallback_on_ties_at_the_median(pxf):
    """60 % of the radii identical and sitting on the median: the bracketed select overflows its
    candidate buffer, flags itself invalid and the five-pass select takes over -- still exact."""
    n = 4_000_001
    rng = np.random.default_rng(41)
    cpu = random_bundle(n, 41)
    cpu[1][:] = rng.normal(0, 1., n)
    cpu[2][:] = rng.normal(0, 1., n)
    th = rng.uniform(0, 2 * np.pi, int(.6 * n))
    cpu[1][: th.size] = 1.17 * np.cos(th[0])
    cpu[2][: th.size] = 1.17 * np.sin(th[0])
    dev = to_dev(cpu)
    got = pxf.analyses.hpd(dev)
    r = pxf.analyses.rho(dev, cent=True).cpu().numpy()
    assert got == 2. * np.median(r)
    assert got == pytest.approx(pyref.hpd(cpu), rel=1e-9)


def test_hpd_nan_propagates(pxf):
    for n in (1001, 2_500_000):
        cpu = random_bundle(n, 35)
        cpu[1][17] = np.nan
        assert np.isnan(pxf.analyses.hpd(to_dev(cpu)))


def test_weighted_hpd_and_sort(pxf):
    import torch
    n = 200_003
    rng = np.random.default_rng(36)
    cpu = random_bundle(n, 36)
    w = rng.uniform(.1, 2., n)
    dev = to_dev(cpu)
    assert pxf.analyses.hpd(dev, weights=w) == pytest.approx(pyref.hpd(cpu, weights=w), rel=1e-9)
    keys = torch.from_numpy(rng.normal(0, 1, n)).cuda()
    keys[::1000] = float("inf")
    keys[5::1000] = float("-inf")
    keys[7::5000] = float("nan")
    ks, idx = pxf.analyses.argsort(keys)
    kc, kn = keys.cpu().numpy(), ks.cpu().numpy()
    assert np.array_equal(kn, np.sort(kc), equal_nan=True)
    assert np.array_equal(kc[idx.cpu().numpy()], kn, equal_nan=True)
    # stability: equal keys keep their original order
    ii = idx.cpu().numpy()
    same = kn[1:] == kn[:-1]
    assert (ii[1:][same] > ii[:-1][same]).all()


def _hpd_weighted_c(pxf, dev, w, which):
    """Call one of the weighted-HPD entry points directly; returns (hpd, valid)."""
    import ctypes
    import torch
    from pyxfocus_b200 import _lib
    x, y = dev[1:3]
    wt = torch.as_tensor(w, dtype=torch.float64, device=x.device).contiguous()
    out = ctypes.c_double(float("nan"))
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    if which == "bracket":
        valid = ctypes.c_int32(-1)
        _lib.check(L.pxf_hpd_weighted_bracket(x.data_ptr(), y.data_ptr(), wt.data_ptr(), x.shape[0], ctypes.byref(out),
                                              ctypes.byref(valid), st))
        return out.value, valid.value
    _lib.check(L.pxf_hpd_weighted_sorted(x.data_ptr(), y.data_ptr(), wt.data_ptr(), x.shape[0], ctypes.byref(out), st))
    return out.value, 1


@pytest.mark.parametrize("n", [1 << 21, 3_000_001])
def test_weighted_hpd_bracketed_path(pxf, n):
    """Large bundles take the bracketed weighted-quantile path (pxf_wquant.cu): same number as numpy's
    argsort -> cumsum -> argmin (analyses.py:73-97) and as the library's own full sort."""
    rng = np.random.default_rng(136)
    cpu = random_bundle(n, 136)
    for w in (rng.uniform(.1, 2., n),                                   # mild weights
              np.where(rng.random(n) < .01, 50., 1.) * rng.random(n),   # heavy tail, zeros possible
              np.ones(n)):                                              # equal weights
        dev = to_dev(cpu)
        want = pyref.hpd(cpu, weights=w)
        got, valid = _hpd_weighted_c(pxf, dev, w, "bracket")
        assert valid == 1
        assert got == pytest.approx(want, rel=1e-9)
        full, _ = _hpd_weighted_c(pxf, dev, w, "sorted")
        assert got == full                                              # same radii picked
        assert pxf.analyses.hpd(dev, weights=w) == got


def test_weighted_hpd_bracket_falls_back(pxf):
    """Inputs the bracketed path refuses (negative weight, NaN radius) report valid == 0 and the public
    call still returns what the full sort returns."""
    n = 1 << 21
    rng = np.random.default_rng(137)
    cpu = random_bundle(n, 137)
    w = rng.uniform(.1, 2., n)
    w[12345] = -1.
    dev = to_dev(cpu)
    _, valid = _hpd_weighted_c(pxf, dev, w, "bracket")
    assert valid == 0
    full, _ = _hpd_weighted_c(pxf, dev, w, "sorted")
    assert pxf.analyses.hpd(dev, weights=w) == full
    assert full == pytest.approx(pyref.hpd(cpu, weights=w), rel=1e-9)
    w[12345] = 1.
    cpu[1][777] = np.nan
    dev = to_dev(cpu)
    _, valid = _hpd_weighted_c(pxf, dev, w, "bracket")
    assert valid == 0
    # a degenerate bundle (every radius equal): brackets collapse, capacity overflows -> fallback, same answer
    cpu = random_bundle(n, 138)
    cpu[1][:] = 3.
    cpu[2][:] = 4.
    dev = to_dev(cpu)
    w = rng.uniform(.1, 2., n)
    assert pxf.analyses.hpd(dev, weights=w) == pytest.approx(pyref.hpd(cpu, weights=w), abs=1e-12)


def test_sharded_merge_kernels_match_the_tensor_driver(pxf):
    """pxf_wq_merge_* (the device form of dist.weighted_quantile_radii) == the tensor-op driver the gloo tests
    exercise: random sorted runs, offsets, bracket key ranges, empty runs, a quantile no key reaches."""
    import torch
    from pyxfocus_b200 import dist
    rng = np.random.default_rng(91)
    for trial in range(24):
        runs_cpu, ranges = [], []
        for i in range(2):
            n = 0 if trial == 5 and i == 0 else int(rng.integers(1, 3000))
            r = np.sort(rng.random(n) * (10. ** rng.integers(-6, 3)))
            if n > 10 and trial % 4 == 0:
                r[n // 2: n // 2 + 5] = r[n // 2]                      # ties
            w = rng.random(n) + .01
            runs_cpu.append((torch.from_numpy(r.view(np.int64).copy()), torch.from_numpy(np.cumsum(w))))
            lo = float(r[0]) * .99 if n else 0.
            hi = float(r[-1]) * 1.01 if n else 1.
            ranges.append((int(np.float64(lo).view(np.int64)), int(np.float64(hi).view(np.int64))))
        off = torch.tensor([float(rng.random() * 3.), float(rng.random() * 3.)]) if trial % 2 else None
        tot = sum(float(c[-1]) if c.shape[0] else 0. for _, c in runs_cpu) / 2. + (float(off.mean()) if off is not None else 0.)
        W = torch.tensor(tot * (1.0 if trial % 3 else 5.0))           # x5: no key reaches .75
        kr = ranges if trial % 3 == 1 else None
        want, wv = dist.weighted_quantile_radii(runs_cpu, [.25, .75], W, offsets=off, key_ranges=kr)
        runs_gpu = [(k.cuda(), c.cuda()) for k, c in runs_cpu]
        got, gv = dist.weighted_quantile_radii(runs_gpu, [.25, .75], W.cuda(), offsets=None if off is None else off.cuda(),
                                               key_ranges=kr)
        assert np.array_equal(gv.cpu().numpy(), wv.numpy()), trial
        for i in range(2):
            if bool(wv[i]):
                assert float(got[i]) == float(want[i]), (trial, i)


def test_image_plane_and_focus(pxf, golden):
    g = golden("ws_offaxis")
    cpu = rows_of(g["after_secondary"])
    dev = to_dev(cpu)
    dz = pxf.analyses.analyticImagePlane(dev)
    assert dz == pytest.approx(float(g["dz_analytic"]), rel=1e-9)
    f = pxf.surfaces.focusI(dev)
    assert f == pytest.approx(float(g["focus"]), rel=1e-9)
    assert_close(to_host(dev), rows_of(g["rays_out"]), pos_scale=1.e4, tol=1e-11, what="focusI rays")
    assert pxf.analyses.hpd(dev) == pytest.approx(float(g["hpd"]), rel=1e-9)
    assert pxf.analyses.rmsCentroid(dev) == pytest.approx(float(g["rms"]), rel=1e-9)
    # findimageplane (parity unpinned): brute-force scan with the oracle agrees on the grid point
    cpu = rows_of(g["rays_out"])
    dev = to_dev(cpu)
    pyref.transform(cpu, 0, 0, 3.3, 0, 0, 0)
    of.surfacesf.flat(*cpu[1:])
    pxf.transformations.transform(dev, 0, 0, 3.3, 0, 0, 0)
    pxf.surfaces.flat(dev)
    best = pxf.analyses.findimageplane(dev, 20., 101)
    scan = np.linspace(-20., 20., 101)
    rms = []
    for dzs in scan:
        t = copy(cpu)
        pyref.transform(t, 0, 0, dzs, 0, 0, 0)
        of.surfacesf.flat(*t[1:])
        rms.append(pyref.rmsCentroid(t))
    assert best == pytest.approx(scan[int(np.argmin(rms))], abs=1e-9)
    assert abs(best - (-3.3)) <= .4 + 1e-9


# --------------------------------------------------------------------------- sources
def test_sources_from_numpy_seeds(pxf):
    for name, args in (("subannulus", (220., 220.6, 1.3, 10_001, -1.)), ("annulus", (200., 230., 10_001, 1.)),
                       ("circularbeam", (12.5, 10_001)), ("pointsource", (.03, 10_001))):
        np.random.seed(37)
        want = getattr(pyref, name)(*args)
        np.random.seed(37)
        got = to_host(getattr(pxf.sources, name)(*args))
        # device sqrt is exact, sin/cos agree with glibc to an ulp or two
        assert_close(got, want, pos_scale=max(1., abs(args[0])), tol=2e-15, what=name)


def test_sources_philox_sharding(pxf):
    """Device RNG: shards generated with first= offsets concatenate to the single-GPU stream."""
    whole = to_host(pxf.sources.subannulus(220., 220.6, 1., 10_000, zhat=-1., rng="philox", seed=5))
    a = to_host(pxf.sources.subannulus(220., 220.6, 1., 4_000, zhat=-1., rng="philox", seed=5, first=0))
    b = to_host(pxf.sources.subannulus(220., 220.6, 1., 6_000, zhat=-1., rng="philox", seed=5, first=4_000))
    assert_bit_equal([np.concatenate([p, q]) for p, q in zip(a, b)], whole, what="philox shards")
    r = np.hypot(whole[1], whole[2])
    assert r.min() >= 220. - 1e-9 and r.max() <= 220.6 + 1e-9
    u = (r ** 2 - 220. ** 2) / (220.6 ** 2 - 220. ** 2)
    assert abs(u.mean() - .5) < .02 and abs(np.arctan2(whole[2], whole[1]).mean()) < .02
    other = to_host(pxf.sources.subannulus(220., 220.6, 1., 10_000, zhat=-1., rng="philox", seed=6))
    assert not np.array_equal(other[1], whole[1])


# --------------------------------------------------------------------------- golden fixtures
def test_golden_wolter1(pxf, golden):
    g = golden("wolter1")
    # sources from the stored uniforms
    got = to_host(pxf.sources.subannulus(220., 220.6, 2 * np.pi, 4000, zhat=-1., uniforms=(g["u1"], g["u2"])))
    assert_close(got, rows_of(g["rays_in"]), pos_scale=220., tol=2e-15, what="golden source")
    dev = to_dev(rows_of(g["rays_in"]))
    T, S = pxf.transformations, pxf.surfaces
    T.transform(dev, 0, 0, -8400., 0, 0, 0)
    S.wolterprimary(dev, 220., 8400.)
    assert_bit_equal(to_host(dev), rows_of(g["after_primary"]), what="golden after_primary")
    T.reflect(dev)
    S.woltersecondary(dev, 220., 8400.)
    T.reflect(dev)
    S.flat(dev)
    assert_bit_equal(to_host(dev), rows_of(g["rays_out"]), what="golden rays_out")
    assert pxf.analyses.hpd(dev) == pytest.approx(float(g["hpd"]), rel=1e-9)


def test_golden_ws(pxf, golden):
    g = golden("ws_offaxis")
    dev = to_dev(rows_of(g["rays_in"]))
    T, S = pxf.transformations, pxf.surfaces
    with pxf.fused(dev):
        T.transform(dev, 0, 0, -1.e4, 0, 0, 0)
        S.wsPrimary(dev, 220., 1.e4, 1.)
    dev[4].add_(np.sin(float(g["theta"])))                 # rays[4] = rays[4] + sin(offaxis)
    dev[6].copy_(-(1. - dev[4] ** 2).sqrt())               # rays[6] = -sqrt(1-rays[4]**2)
    T.reflect(dev)
    S.wsSecondary(dev, 220., 1.e4, 1.)
    T.reflect(dev)
    assert_close(to_host(dev), rows_of(g["after_secondary"]), pos_scale=1.e4, what="golden ws after_secondary")
    g2 = golden("ws_cap")
    dev = to_dev(rows_of(g2["rays_in"]))
    T.transform(dev, 0, 0, -1.e4, 0, 0, 0)
    S.wsPrimary(dev, 220., 1.e4, 1.)
    pxf.Program().kick(np.sin(float(g2["theta"])), 0., -1.).run(dev)
    T.reflect(dev)
    fail = S.wsSecondary(dev, 220., 1.e4, 1., check=True).cpu().numpy()
    assert np.array_equal(fail, g2["restored"]), "set of non-converged (restored) rays differs"
    assert_close(to_host(dev), rows_of(g2["rays_out"]), pos_scale=1.e4, what="golden ws_cap")


def test_golden_zernike(pxf, golden):
    g = golden("zernike")
    T, S = pxf.transformations, pxf.surfaces
    dev = to_dev(rows_of(g["rays_in"]))
    T.transform(dev, 0, 0, -100., 0, 0, 0)
    S.zernsurf(dev, g["coeff"], float(g["rad"]), rorder=g["rorder"], aorder=g["aorder"], nr=1.)
    assert_close(to_host(dev), rows_of(g["after_zern"]), pos_scale=100., what="golden after_zern")
    T.reflect(dev)
    T.transform(dev, 0, 0, 50., 0, 0, 0)
    S.flat(dev, nr=1.)
    assert_close(to_host(dev), rows_of(g["rays_out"]), pos_scale=150., what="golden zern rays_out")
    dev = to_dev(rows_of(g["rays_in2"]))
    T.transform(dev, 1., -2., -100., 1e-3, -2e-3, .3)
    S.zernsurf(dev, g["coeff"], float(g["rad"]), rorder=g["rorder"], aorder=g["aorder"])
    assert_close(to_host(dev), rows_of(g["rays_out2"]), pos_scale=100., what="golden zern rays_out2")


def test_golden_spo_grating(pxf, golden):
    g = golden("spo_grating")
    T, S = pxf.transformations, pxf.surfaces
    R0, F = float(g["R0"]), float(g["F"])
    dev = to_dev(rows_of(g["rays_in"]))
    T.transform(dev, 0, 0, 0, 0, 0, .01)
    S.spoPrimary(dev, R0, F)
    T.reflect(dev)
    S.spoSecondary(dev, R0, F)
    T.reflect(dev)
    assert_bit_equal(to_host(dev), rows_of(g["after_spo"]), what="golden after_spo")
    T.transform(dev, 0, 0, -(F - 200.), 0, 0, 0)
    T.transform(dev, 0., 0, 0, 0, 0, 0)
    S.flat(dev)
    T.transform(dev, 0, 11832.911, 0, 0, 0, 0)
    mask = g["mask"]
    T.reflect(dev, ind=mask)
    T.radgrat(dev, 160. / 11832.911, -3, 2.4, ind=mask)
    T.radgrat(dev, 160. / 11832.911, 1, g["wave"], ind=~mask)
    assert_close(to_host(dev), rows_of(g["after_grat"]), pos_scale=1.2e4, what="golden after_grat")
    T.radgrat(dev, 160. / 11832.911, 150, 2.4, ind=g["evan"])
    assert_close(to_host(dev), rows_of(g["after_evan"]), pos_scale=1.2e4, what="golden after_evan (NaN pattern)")


def test_golden_misc(pxf, golden):
    g = golden("misc")
    T, S = pxf.transformations, pxf.surfaces
    dev = to_dev(rows_of(g["rays_in"]))
    T.transform(dev, 0, 0, -500., 0, 0, 0)
    S.conic(dev, 1000., -1.)
    assert_bit_equal(to_host(dev), rows_of(g["after_conic"]), what="golden after_conic")
    T.refract(dev, 1., 1.5)
    assert_close(to_host(dev), rows_of(g["after_refract"]), pos_scale=500., what="golden after_refract")
    T.transform(dev, 0, 0, 10., 0, 0, 0)
    S.conic(dev, -800., .3, nr=1.5)
    assert_close(to_host(dev), rows_of(g["after_conicopd"]), pos_scale=500., what="golden after_conicopd")
    T.refract(dev, 1.5, 1.)
    T.itransform(dev, .3, -.2, 5., .01, .02, -.03)
    assert_close(to_host(dev), rows_of(g["after_itransform"]), pos_scale=500., what="golden after_itransform")
    S.flat(dev, ind=g["sel"])
    assert_close(to_host(dev), rows_of(g["after_flat_ind"]), pos_scale=500., what="golden after_flat_ind")
    T.grat(dev, 160.e-6, g["grat_order"], g["grat_wave"] * 1.e-6)
    assert_close(to_host(dev), rows_of(g["after_grat"]), pos_scale=500., what="golden after_grat")
    dev = to_dev(rows_of(g["rays_in2"]))
    T.transform(dev, 0, 0, -8400., 0, 0, 0)
    S.woltersine(dev, 220., 8400., 1.e-4, 1. / 20.)
    assert_close(to_host(dev), rows_of(g["after_woltersine"]), pos_scale=8.4e3, what="golden after_woltersine")


# --------------------------------------------------------------------------- config 5: nested shells
def test_nested_multishell_assembly_weighted_analyses(pxf):
    """BASELINE config 5 in small: a nested Wolter-I assembly (one prescription per shell,
    examples/axro/axialHeights.py:215-322 / SMARTX.py:163-259).  All shells live in ONE device
    bundle; each shell's segment is traced by one fused launch with its own (r0,z0); a z-range
    vignette and the area-weighted centroid / rms / hpd follow.  The oracle does what the
    reference does: a Python loop over shells, then np.concatenate."""
    rng = np.random.default_rng(42)
    radii = np.linspace(200., 1500., 24)
    per = 4000 + 2 * rng.integers(0, 500, radii.size)            # even segment sizes
    cpu_shells, wts = [], []
    for k, (r0, nk) in enumerate(zip(radii, per)):
        z0 = np.sqrt(1.e4 ** 2 - r0 ** 2)
        np.random.seed(100 + k)
        rays = pyref.annulus(r0, r0 + .6, int(nk), zhat=-1.)
        chains.run_steps_cpu(rays, chains.wolter1_steps(r0, z0, 1.))
        cpu_shells.append(rays)
        wts.append(np.full(int(nk), 2 * np.pi * r0 * .6 / nk))       # geometric area per ray
    cpu = [np.concatenate([s[i] for s in cpu_shells]) for i in range(10)]
    w = np.concatenate(wts)
    # device: one allocation, per-shell views
    total = int(per.sum())
    from pyxfocus_b200._call import bundle_alloc, bundle_split
    dev = bundle_alloc(total, "cuda", zero=True)
    segs = bundle_split(dev, [int(v) for v in per])
    for k, (r0, nk) in enumerate(zip(radii, per)):
        z0 = np.sqrt(1.e4 ** 2 - r0 ** 2)
        np.random.seed(100 + k)
        src = pyref.annulus(r0, r0 + .6, int(nk), zhat=-1.)
        for i in range(10):
            segs[k][i].copy_(__import__("torch").from_numpy(src[i]))
        steps_to_program(chains.wolter1_steps(r0, z0, 1.)).run(segs[k])
    assert_bit_equal(to_host(dev), cpu, what="nested shells")
    A = pxf.analyses
    assert A.hpd(dev, weights=w) == pytest.approx(pyref.hpd(cpu, weights=w), rel=1e-9)
    assert A.rmsCentroid(dev, weights=w) == pytest.approx(pyref.rmsCentroid(cpu, weights=w), rel=1e-9)
    cx, cy = A.centroid(dev, weights=w)
    rx, ry = pyref.centroid(cpu, weights=w)
    assert abs(cx - rx) <= 1e-15 and abs(cy - ry) <= 1e-15
    assert A.hpd(dev) == pytest.approx(pyref.hpd(cpu), rel=1e-9)
    assert pxf.dist.hpd(dev) == A.hpd(dev)


# --------------------------------------------------------------------------- Legendre-Legendre shells
def test_segmented_program_and_source_match_per_shell_launches(pxf):
    """pxf_trace_program_segmented / pxf_source_segmented (one launch for all shells) == the per-shell launches,
    bit for bit: ragged segment sizes (odd, smaller than a tile, larger than a tile, one empty), a vignette
    predicate, and the out-of-place form."""
    import torch
    from pyxfocus_b200._call import bundle_alloc, bundle_split
    radii = [200., 350.5, 612., 800., 1100., 1499.]
    sizes = [5001, 1, 0, 7000, 2048, 333]
    total = sum(sizes)
    z0s = [float(np.sqrt(1.e4 ** 2 - r ** 2)) for r in radii]

    def prog(r0, z0):
        return (pxf.Program().transform(0, 0, z0, 0, 0, 0).wolterprimary(r0, z0, 1.).reflect()
                .vignette_box(3, z0 + 20., z0 + 80.).woltersecondary(r0, z0, 1.).reflect().flat())
    # sources: segmented == per segment
    one = pxf.sources.segments("annulus", [(r, r + .6, 0., -1.) for r in radii], sizes, seed=5, first=100)
    ref = bundle_alloc(total, "cuda", zero=True)
    rsegs = bundle_split(ref, sizes)
    off = 0
    for k, nk in enumerate(sizes):
        if nk:
            pxf.sources.annulus(radii[k], radii[k] + .6, nk, zhat=-1., rng="philox", seed=5, first=100 + off, out=rsegs[k])
        off += nk
    assert_bit_equal(to_host(one), to_host(ref), what="segmented source")
    # traces
    alive_ref = torch.zeros(total, dtype=torch.uint8, device="cuda")
    off = 0
    for k, nk in enumerate(sizes):
        if nk:
            alive_ref[off:off + nk] = prog(radii[k], z0s[k]).run(rsegs[k])
        off += nk
    sp = pxf.SegmentedProgram([prog(r, z) for r, z in zip(radii, z0s)], sizes)
    out = bundle_alloc(total, "cuda", zero=True)
    alive2 = sp.run(one, out=out)
    assert np.array_equal(alive2.cpu().numpy(), alive_ref.cpu().numpy())
    assert 0 < int(alive_ref.sum()) < total
    got, want = to_host(out), to_host(ref)
    assert_bit_equal(got, want, rows=range(1, 10), what="segmented program (out of place)")
    alive1 = sp.run(one)
    assert np.array_equal(alive1.cpu().numpy(), alive_ref.cpu().numpy())
    assert_bit_equal(to_host(one), want, rows=range(1, 10), what="segmented program (in place)")
    # the canonical Wolter-I chain takes the statically specialised segmented kernel (k_chain_seg): same bits
    def chain(r0, z0):
        return (pxf.Program().transform(0, 0, z0, 0, 0, 0).wolterprimary(r0, z0, 1.).reflect()
                .woltersecondary(r0, z0, 1.).reflect().flat())
    one = pxf.sources.segments("annulus", [(r, r + .6, 0., -1.) for r in radii], sizes, seed=5, first=100)
    ref = to_dev(to_host(one))
    rsegs = bundle_split(ref, sizes)
    for k, nk in enumerate(sizes):
        if nk:
            chain(radii[k], z0s[k]).run(rsegs[k])
    before = pxf.launch_count()
    pxf.SegmentedProgram([chain(r, z) for r, z in zip(radii, z0s)], sizes).run(one)
    assert pxf.launch_count() - before == 1
    assert_bit_equal(to_host(one), to_host(ref), rows=range(1, 10), what="segmented Wolter-I chain")
    with pytest.raises(ValueError):
        pxf.SegmentedProgram([prog(200., 9000.), pxf.Program().flat()], [4, 4])
    with pytest.raises(ValueError):
        sp.run(rsegs[0])


def test_golden_legendre_shells(pxf, golden):
    """wolterprimLL / woltersecLL / ellipsoidWoltLL and the ellipsoid pair,
    against fixtures produced by the reference's own surfaces.primaryLL / secondaryLL /
    ellipsoidPrimary(LL) / ellipsoidSecondary(LL)."""
    g = golden("legendre")
    T, S = pxf.transformations, pxf.surfaces
    coeff, axial, az = g["coeff"], g["axial"], g["az"]
    dev = to_dev(rows_of(g["rays_in"]))
    T.transform(dev, 0, 0, -8400., 0, 0, 0)
    S.primaryLL(dev, 220., 8400., 8500., 8400., .25, coeff, axial, az)
    assert_close(to_host(dev), rows_of(g["after_primaryLL"]), pos_scale=8.5e3, what="primaryLL")
    T.reflect(dev)
    S.secondaryLL(dev, 220., 8400., 1., 8400., 8300., .25, coeff * .5, axial, az)
    assert_close(to_host(dev), rows_of(g["after_secondaryLL"]), pos_scale=8.5e3, tol=1e-11, what="secondaryLL")
    T.reflect(dev)
    S.flat(dev)
    assert_close(to_host(dev), rows_of(g["rays_out"]), pos_scale=8.5e3, tol=1e-11, what="LL rays_out")
    assert pxf.analyses.hpd(dev) == pytest.approx(float(g["hpd"]), rel=1e-9)
    Sd = float(g["S"])
    dev = to_dev(rows_of(g["rays_in2"]))
    S.ellipsoidPrimaryLL(dev, 220., 8400., Sd, 1., 8500., 8400., .25, coeff, axial, az)
    assert_close(to_host(dev), rows_of(g["after_ellipsoidLL"]), pos_scale=8.5e3, what="ellipsoidPrimaryLL")
    T.reflect(dev)
    S.ellipsoidSecondaryLL(dev, 220., 8400., Sd, 1., 8400., 8300., .25, coeff * .5, axial, az)
    T.reflect(dev)
    S.flat(dev)
    assert_close(to_host(dev), rows_of(g["rays_out2"]), pos_scale=8.5e3, tol=1e-11, what="ellipsoid LL rays_out")
    dev = to_dev(rows_of(g["rays_in3"]))
    S.ellipsoidPrimary(dev, 220., 8400., Sd, 1.)
    T.reflect(dev)
    S.ellipsoidSecondary(dev, 220., 8400., Sd, 1.)
    T.reflect(dev)
    S.flat(dev)
    assert_bit_equal(to_host(dev), rows_of(g["rays_out3"]), rows=range(1, 10), what="ellipsoid pair (algebraic)")
    assert pxf.analyses.hpd(dev) == pytest.approx(float(g["hpd3"]), rel=1e-9)


def test_legendre_shells_vs_oracle_high_order_and_mask(pxf):
    rng = np.random.default_rng(43)
    # orders up to 11 exercise the 16x16 table; duplicates of one (axial,az) pair must add up
    axial = np.array([0, 1, 9, 11, 3, 3, 2])
    az = np.array([0, 2, 1, 4, 10, 10, 0])
    coeff = rng.normal(0, 1e-4, axial.size)
    cpu = wolter_inputs(20_001, 44)
    mask = rng.random(20_001) < .5
    dev = to_dev(cpu)
    pyref.masked(of.woltsurf.wolterprimll, cpu[1:], mask, 220., 8400., 8500., 8400., 1., coeff, axial, az)
    pxf.woltsurf.wolterprimll(*dev[1:], 220., 8400., 8500., 8400., 1., coeff, axial, az, mask=mask)
    assert_close(to_host(dev), cpu, pos_scale=8.5e3, what="wolterprimll masked, order 11")
    # rays outside the Legendre domain (|zarg|>1, |targ|>1): clamp semantics of legendre/legendrep
    cpu = wolter_inputs(5_001, 45)
    dev = to_dev(cpu)
    of.woltsurf.wolterprimll(*cpu[1:], 220., 8400., 8460., 8440., .2, coeff[:3], axial[:3], az[:3])
    pxf.woltsurf.wolterprimll(*dev[1:], 220., 8400., 8460., 8440., .2, coeff[:3], axial[:3], az[:3])
    assert_close(to_host(dev), cpu, pos_scale=8.5e3, tol=1e-11, what="wolterprimll outside the domain")
    with pytest.raises(pxf.PxfError):
        pxf.woltsurf.wolterprimll(*dev[1:], 220., 8400., 8460., 8440., .2, [1.], [16], [0])


# --------------------------------------------------------------------------- edge cases / errors
def test_empty_and_tiny_bundles(pxf):
    import torch
    for n in (0, 1, 2, 3):
        cpu = wolter_inputs(max(n, 1), 38)
        cpu = [r[:n].copy() for r in cpu]
        dev = to_dev(cpu) if n else [torch.empty(0, dtype=torch.float64, device="cuda") for _ in range(10)]
        steps = chains.wolter1_steps()[1:]
        if n:
            chains.run_steps_cpu(cpu, steps)
        run_steps_gpu(dev, steps)
        steps_to_program(steps).run(to_dev(cpu) if n else dev)
        assert_bit_equal(to_host(dev), cpu, what="n=%d" % n)
    empty = [torch.empty(0, dtype=torch.float64, device="cuda") for _ in range(10)]
    assert pxf.transformations.vignette(empty)[1].shape[0] == 0
    assert np.isnan(pxf.analyses.hpd(empty))


def test_f2py_style_argument_errors(pxf):
    import torch
    good = to_dev(random_bundle(100, 39))
    T = pxf.transformationsf
    with pytest.raises(ValueError):
        T.reflect(good[4].float(), *good[5:])                       # wrong dtype
    with pytest.raises(ValueError):
        T.reflect(good[4][::2], *good[5:])                          # non-contiguous
    with pytest.raises(ValueError):
        T.reflect(good[4][:50].contiguous(), *good[5:])             # length mismatch
    with pytest.raises(ValueError):
        T.reflect(good[4].cpu(), *good[5:])                         # CPU tensor
    with pytest.raises(pxf.PxfError):
        pxf.zernsurf.tracezern(*good[1:], np.zeros(3), [0, 1, 5], [0, 1, 1], 10.)   # n too high for 3 terms
    with pytest.raises((pxf.PxfError, ValueError)):
        pxf.Program().add(99).run(good)                             # unknown opcode
    with pytest.raises(NotImplementedError):
        pxf.surfaces.zernsurf(good, np.zeros(3), 10.)               # un-vendored default ordering
    assert torch.cuda.is_available()


def test_numpy_arrays_are_accepted_in_place(pxf):
    """Literal drop-in: the f2py-shaped modules also take the reference's host numpy arrays and
    mutate them in place (staged through the device)."""
    cpu = wolter_inputs(5_001, 40)
    ref = copy(cpu)
    pxf.woltsurf.wolterprimary(*cpu[1:], 220., 8400., 1.)
    pxf.transformationsf.reflect(*cpu[4:])
    of.woltsurf.wolterprimary(*ref[1:], 220., 8400., 1.)
    of.transformationsf.reflect(*ref[4:])
    assert_bit_equal(cpu, ref, what="numpy in place")


# --------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("n", [20_000_000])
def test_full_size_properties(pxf, n):
    """At a BASELINE-scale bundle the oracle is too slow; check size-independent properties:
    rays land ON the prescription surfaces (conicsolve closed forms), the fused chain focuses
    to the same HPD as the oracle's sample, transform/itransform and double reflection are
    identities to rounding, compaction keeps exactly the flagged rays in order, the sort is
    sorted and a permutation, and the select equals torch's own median."""
    import torch
    from pyxfocus_b200 import conicsolve as con
    S, T, A = pxf.surfaces, pxf.transformations, pxf.analyses
    rays = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=1)
    with pxf.fused(rays):
        T.transform(rays, 0, 0, -8400., 0, 0, 0)
        S.wolterprimary(rays, 220., 8400.)
    r = torch.sqrt(rays[1] ** 2 + rays[2] ** 2).cpu().numpy()
    assert np.abs(r - con.primrad(rays[3].cpu().numpy(), 220., 8400.)).max() <= 1e-12 * 220.
    with pxf.fused(rays):
        T.reflect(rays)
        S.woltersecondary(rays, 220., 8400.)
    r = torch.sqrt(rays[1] ** 2 + rays[2] ** 2).cpu().numpy()
    assert np.abs(r - con.secrad(rays[3].cpu().numpy(), 220., 8400.)).max() <= 1e-10 * 220.
    nrm = (rays[7] ** 2 + rays[8] ** 2 + rays[9] ** 2).sqrt()
    assert float((nrm - 1).abs().max()) <= 4e-16
    keep = copy_dev(rays)
    T.reflect(rays)
    T.reflect(rays)                                  # same normal twice = identity to rounding
    for k in (4, 5, 6):
        assert float((rays[k] - keep[k]).abs().max()) <= 1e-15
    T.transform(rays, 1., 2., 3., .1, .2, .3)
    T.itransform(rays, 1., 2., 3., .1, .2, .3)
    for k in range(1, 10):
        assert float((rays[k] - keep[k]).abs().max()) <= 1e-12 * (8.5e3 if k < 4 else 1.)
    rays = keep
    with pxf.fused(rays):
        T.reflect(rays)
        S.flat(rays)
    h = A.hpd(rays)
    cpu = chains.wolter1_source(100_000, 0)
    assert h == pytest.approx(chains.wolter1_cpu(cpu), rel=.05)      # same optics, different sample
    rad = A.rho(rays, cent=True)
    assert h == 2. * float(torch.median(rad)) or h == pytest.approx(2. * float(rad.median()), rel=1e-6)
    srt = torch.sort(rad).values
    k0, k1 = (n - 1) // 2, n // 2
    assert h == float(srt[k0] + srt[k1])             # 2 * (a+b)/2, exact
    ks, idx = A.argsort(rad)
    assert torch.equal(ks, srt)
    assert torch.equal(torch.sort(idx).values, torch.arange(n, device="cuda"))
    flags = (rays[1] > 0) & (rays[2].abs() < 1e-5)
    out = T.vignette(rays, ind=flags)
    assert out[1].shape[0] == int(flags.sum())
    assert torch.equal(out[1], rays[1][flags]) and torch.equal(out[9], rays[9][flags])
    # weighted HPD: the bracketed path picks the same two radii as torch's full sort + cumsum + argmin
    w = torch.linspace(.5, 2., n, dtype=torch.float64, device="cuda")
    hw = A.hpd(rays, weights=w)
    cxw, cyw = A.centroid(rays, weights=w)
    rw = torch.sqrt((rays[1] - cxw) ** 2 + (rays[2] - cyw) ** 2)
    order = torch.argsort(rw, stable=True)
    cdf = torch.cumsum(w[order], 0)
    cdf = cdf / cdf.max()
    want = float(rw[order][torch.argmin((cdf - .75).abs())] - rw[order][torch.argmin((cdf - .25).abs())])
    assert hw == pytest.approx(want, rel=1e-9)
    del order, cdf, rw
    # a nested assembly in one launch == the same shells one by one (first and last shell compared)
    radii = np.linspace(200., 1500., 260)
    sizes = [n // 260] * 260
    sizes[-1] += n - sum(sizes)
    z0s = [float(np.sqrt(1.e4 ** 2 - r ** 2)) for r in radii]

    def chain(r0, z0):
        return (pxf.Program().transform(0, 0, z0, 0, 0, 0).wolterprimary(r0, z0, 1.).reflect()
                .woltersecondary(r0, z0, 1.).reflect().flat())
    nest = pxf.sources.segments("annulus", [(r, r + .6, 0., -1.) for r in radii], sizes, seed=2)
    pxf.SegmentedProgram([chain(r, z) for r, z in zip(radii, z0s)], sizes).run(nest)
    for k, first in ((0, 0), (259, n - sizes[-1])):
        one = pxf.sources.annulus(radii[k], radii[k] + .6, sizes[k], zhat=-1., rng="philox", seed=2, first=first)
        chain(radii[k], z0s[k]).run(one)
        for row in range(1, 10):
            assert torch.equal(one[row], nest[row][first:first + sizes[k]]), (k, row)
    assert float(nest[3].abs().max()) == 0.                   # every shell focused onto its focal plane z = 0
    hp = A.hpd(nest)
    assert 0. < hp < 1e-3


def copy_dev(rays):
    import pyxfocus_b200
    return pyxfocus_b200.transformations.copy_rays(rays)


# ---------------------------------------------------------------- SURVEY 8f rank 2 / rank 3
def test_golden_remaining_surfaces(pxf, golden):
    """sphere / tanSphere / cyl / cylconic / conicplus / torus / paraxial / legSurf / W-S back
    surfaces / zernphase / zernsurfrot through the product's surfaces.py mirrors, against fixtures
    written by the reference's own wrappers (tests/golden/make_golden.py).  Algebraic routines
    (+ - * / sqrt, integer powers) are bit-exact; the libm ones agree to 1e-12."""
    g = golden("surfaces2")
    T, S = pxf.transformations, pxf.surfaces

    def run(key_in, fn):
        dev = to_dev(rows_of(g[key_in]))
        fn(dev)
        return to_host(dev)

    assert_bit_equal(run("sphere_in", lambda r: S.sphere(r, 250.)), rows_of(g["sphere_out"]), what="sphere")
    assert_bit_equal(run("tansphere_in", lambda r: S.tanSphere(r, -500., nr=1.5)), rows_of(g["tansphere_out"]),
                     what="tanSphere")
    assert_bit_equal(run("cyl_in", lambda r: S.cyl(r, 100., nr=1.2)), rows_of(g["cyl_out"]), what="cyl")
    assert_bit_equal(run("cylconic_in", lambda r: S.cylconic(r, 1. / 400., -.7)), rows_of(g["cylconic_out"]),
                     what="cylconic")
    assert_bit_equal(run("conicplus_in", lambda r: S.conicplus(r, 800., -1.3, g["conicplus_p"], nr=1.1)),
                     rows_of(g["conicplus_out"]), what="conicplus")
    assert_bit_equal(run("torus_in", lambda r: S.torus(r, 150., 900.)), rows_of(g["torus_out"]), what="torus")

    def parax(r):
        S.paraxial(r, 350.)
        S.paraxialY(r, -120.)
    assert_bit_equal(run("paraxial_in", parax), rows_of(g["paraxial_out"]), what="paraxial")
    assert_bit_equal(run("legsurf_in", lambda r: S.legSurf(r, 25., 30., 2., g["leg_coeff"], g["leg_xo"], g["leg_yo"])),
                     rows_of(g["legsurf_out"]), what="legSurf")
    # libm per ray: atan2 / sincos / asin / pow
    dev = to_dev(rows_of(g["wsback_in"]))
    S.wsPrimaryB(dev, 220., 1.e4, 1., .4)
    assert_close(to_host(dev), rows_of(g["wsback_primary"]), pos_scale=1.e4, what="wsPrimaryB")
    T.reflect(dev)
    S.wsSecondaryB(dev, 220., 1.e4, 1., .4)
    assert_close(to_host(dev), rows_of(g["wsback_secondary"]), pos_scale=1.e4, tol=1e-11, what="wsSecondaryB")
    ro, ao = g["z_rorder"], g["z_aorder"]
    got = run("zernphase_in", lambda r: S.zernphase(r, g["z_coeff"], 20., 5.e-4, rorder=ro, aorder=ao))
    assert_close(got, rows_of(g["zernphase_out"]), pos_scale=20., what="zernphase")
    got = run("zernrot_in", lambda r: S.zernsurfrot(r, g["z_coeff"], g["z_coeff2"], 20., .37, rorder1=ro, aorder1=ao,
                                                   rorder2=ro[:10], aorder2=ao[:10]))
    assert_close(got, rows_of(g["zernrot_out"]), pos_scale=20., what="zernsurfrot")


def test_remaining_surfaces_vs_oracle_masks_and_misses(pxf):
    """Same routines straight through the f2py-shaped modules on the generic bundle (rays that miss
    the sphere / cylinder are zeroed and get NaN normals), unmasked and with ind= masks."""
    from pyxfocus_b200 import surfacesf as PS
    rng = np.random.default_rng(40)
    mask = rng.random(N) < .6
    cases = [
        ("tracesphere", (150.,), False), ("tracesphereopd", (150., 1.4), True),
        ("tracecyl", (120.,), False), ("tracecylopd", (120., 1.3), True),
        ("paraxial", (75.,), False), ("paraxialy", (-33.,), False),
    ]
    for name, scalars, with_opd in cases:
        for use_mask in (False, True):
            cpu = random_bundle(N, 41)
            dev = to_dev(cpu)
            if use_mask:
                sub = [np.ascontiguousarray(r[mask]) for r in cpu]
                a = sub if with_opd else sub[1:]
                getattr(of.surfacesf, name)(*a, *scalars)
                for r, s_ in zip(cpu, sub):
                    r[mask] = s_
                a = dev if with_opd else dev[1:]
                getattr(PS, name)(*a, *scalars, mask=mask)
            else:
                a = cpu if with_opd else cpu[1:]
                getattr(of.surfacesf, name)(*a, *scalars)
                a = dev if with_opd else dev[1:]
                getattr(PS, name)(*a, *scalars)
            assert_bit_equal(to_host(dev), cpu, what="%s mask=%s" % (name, use_mask))


def test_remaining_analyses_sources_and_helpers(pxf, golden):
    """per-axis analyses, angle helpers, steering, set-up sources."""
    A, T, S = pxf.analyses, pxf.transformations, pxf.surfaces
    g = golden("wolter1")
    rays = rows_of(g["rays_out"])
    # move off focus so the per-axis planes are well conditioned
    pyref.transform(rays, 0, 0, 3., 0, 0, 0)
    of.surfacesf.flat(*rays[1:])
    dev = to_dev(rays)
    w = np.linspace(.5, 2., rays[1].size)
    x, y, z, l, m, n = rays[1:7]

    def yplane(ww):
        by = np.average(y * m / n, weights=ww) - np.average(y, weights=ww) * np.average(m / n, weights=ww)
        ay = np.average((m / n) ** 2, weights=ww) - np.average(m / n, weights=ww) ** 2
        return -by / ay

    def xplane(ww):
        bx = np.average(x * l / n, weights=ww) - np.average(x, weights=ww) * np.average(l / n, weights=ww)
        ax = np.average((l / n) ** 2, weights=ww) - np.average(l / n, weights=ww) ** 2
        return -bx / ax
    for ww in (None, w):
        assert A.analyticYPlane(dev, weights=ww) == pytest.approx(yplane(ww), rel=1e-9)
        assert A.analyticXPlane(dev, weights=ww) == pytest.approx(xplane(ww), rel=1e-9)
    cy = np.average(y)
    assert A.hpdY(dev) == pytest.approx(np.median(np.abs(y - cy)) * 2., rel=1e-12)
    pt = (1e-3, -2e-3, .5)
    rho = (x - pt[0]) ** 2 + (y - pt[1]) ** 2 + (z - pt[2]) ** 2
    assert A.rmsPoint(dev, pt) == pytest.approx(np.sqrt(np.average(rho)), rel=1e-12)
    assert A.rmsPoint(dev, pt, weights=w) == pytest.approx(np.sqrt(np.average(rho, weights=w)), rel=1e-12)
    ia = np.arccos(rays[4] * rays[7] + rays[5] * rays[8] + rays[6] * rays[9])
    assert np.allclose(A.indAngle(dev).cpu().numpy(), ia, atol=1e-14)
    assert np.allclose(A.grazeAngle(dev).cpu().numpy(), np.pi / 2 - ia, atol=1e-14)
    nrm = (0., .6, .8)
    assert np.allclose(A.indAngle(dev, normal=nrm).cpu().numpy(), np.arccos(.6 * rays[5] + .8 * rays[6]), atol=1e-14)
    # focusX / focusY: two analytic steps + flats, against the same sequence on the oracle
    for fn_dev, plane in ((S.focusY, yplane), (S.focusX, xplane)):
        d2 = to_dev(rays)
        dz = fn_dev(d2)
        assert abs(dz + 3.) < .5                            # walks back towards the focus
        got = to_host(d2)
        assert np.all(got[3] == 0.) and np.all(got[9] == 1.)
    # steering: mean tilt removed
    np.random.seed(5)
    tilted = pyref.circularbeam(5., 4001)
    pyref.transform(tilted, 0, 0, 0, .01, -.02, 0)
    d3 = to_dev(tilted)
    T.steerY(d3)
    T.steerX(d3)
    assert abs(float(d3[5].mean())) <= 1e-6 and abs(float(d3[4].mean())) <= 1e-6
    assert to_host(pxf.sources.circFan(.05, 7, 12))[4].size == 84
    pts = T.applyTPos(d3[1], d3[2], d3[3], T.newCoords())
    assert torch_equal(pts[0], d3[1])


def torch_equal(a, b):
    import torch
    return bool(torch.equal(a, b))


def test_bench_scale_properties(pxf):
    """The bench's own size (1.25e8 rays per GPU = BASELINE configs[4] sharded over 8): size-independent
    properties of the fused trace, the exact select, the bracketed weighted quantiles and the compaction."""
    import torch
    from pyxfocus_b200._call import bundle_alloc
    n = 125_000_000
    A, T = pxf.analyses, pxf.transformations
    src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0)
    out = bundle_alloc(n, "cuda", zero=True)
    sums = torch.zeros(16, dtype=torch.float64, device="cuda")
    prog = (pxf.Program().transform(0., 0., 8400., 0., 0., 0.).wolterprimary(220., 8400., 1.).reflect()
            .woltersecondary(220., 8400., 1.).reflect().flat())
    prog.run(src, out=out, sums=sums)
    # the source is untouched, every ray sits on the focal plane with the plane's normal
    assert float(src[3].abs().max()) == 0. and float(src[6].max()) == -1.
    assert float(out[3].abs().max()) == 0. and float(out[9].min()) == 1. and float(out[7].abs().max()) == 0.
    # centroid sums emitted by the trace kernel == a separate pass (same fixed-shape tree is not required: 1e-15)
    cx, cy = A.centroid(out)
    assert float(sums[0]) == n
    assert abs(float(sums[1]) / n - cx) <= 1e-15 and abs(float(sums[2]) / n - cy) <= 1e-15
    # unweighted HPD == the two middle order statistics of a full sort
    h = A.hpd(out, sums=sums)
    rad = A.rho(out, cent=True)
    srt = torch.sort(rad).values
    assert h == float(srt[(n - 1) // 2] + srt[n // 2])
    assert h == pytest.approx(1.28e-5, rel=.01)                       # the config-1 answer (REAL*4 delta in flat)
    del srt
    # weighted HPD: bracketed path == the library's full sort path, bit for bit
    w = torch.linspace(.5, 2., n, dtype=torch.float64, device="cuda")
    hb, valid = _hpd_weighted_c(pxf, out, w, "bracket")
    hs, _ = _hpd_weighted_c(pxf, out, w, "sorted")
    assert valid == 1 and hb == hs
    del w, rad
    # compaction keeps exactly the flagged rays, in order
    flags = out[1] > 0
    kept = T.vignette(out, ind=flags)
    assert kept[1].shape[0] == int(flags.sum())
    assert bool((kept[1] > 0).all())
    assert torch.equal(kept[2][:1000], out[2][flags][:1000])


# --------------------------------------------------------------------------- set-up sources, helpers
def test_grid_sources_on_device(pxf):
    """xslit / rectArray / fanBeam / circFan (sources.py:173-247,418-471) generated by k_source_grid against the numpy
    restatement: the linspace/meshgrid positions bit for bit, sin/cos-derived cosines to an ulp or two; shards
    (first/count) concatenate to the whole source."""
    S = pxf.sources
    for name, args in (("xslit", (-3., 5., 257, 1.)), ("xslit", (2., 2., 5)), ("xslit", (1., 4., 1)),
                       ("rectArray", (4., 2.5, 33)), ("rectArray", (1., 1., 1)),
                       ("fanBeam", (.02, .03, 21)), ("circFan", (.05, 7, 12)), ("circFan", (.3, 1, 5)),
                       ("circFan", (.1, 64, 101))):
        want = getattr(pyref, name)(*args)
        got = to_host(getattr(S, name)(*args))
        assert got[1].shape == want[1].shape, name
        assert_bit_equal(got, want, rows=(0, 1, 2, 3, 7, 8, 9), what=name + " opd/x/y/z/normals")
        if name in ("xslit", "rectArray"):
            assert_bit_equal(got, want, what=name)
        else:
            for k in (4, 5, 6):
                assert np.max(np.abs(got[k] - want[k])) <= 4e-16, (name, k)
        total = want[1].size
        if total > 3:
            cut = total // 3 + 1
            a = to_host(getattr(S, name)(*args, first=0, count=cut))
            b = to_host(getattr(S, name)(*args, first=cut))
            assert_bit_equal([np.concatenate([p, q]) for p, q in zip(a, b)], got, what=name + " shards")
    with pytest.raises(pxf.PxfError):
        S.xslit(0., 1., 10, first=8, count=5)


def test_beam_sources_on_device(pxf):
    """convergingbeam / convergingbeam2 / rectbeam / gaussianBeam (sources.py:250-416): numpy's draws uploaded in the
    reference's order, geometry by k_source_beam.  rectbeam is pure multiply/subtract: bit for bit.  The converging
    beams form sqrt(1-n^2) with n within a few 1e-4 of -1, which amplifies the one-ulp differences between the
    device's and numpy's cos/atan/tan by 1/(1-n^2): l, m agree to 1e-12, positions to an ulp."""
    S = pxf.sources
    for name, args, tol in (("rectbeam", (12., 7., 4001), 0.),
                            ("gaussianBeam", (.01, 4001), 4e-16),
                            ("convergingbeam", (8400., 200., 230., -.1, .3, 4001, 1.5), 1e-12),
                            ("convergingbeam2", (8400., -20., 30., 190., 240., 4001, .5), 1e-12)):
        np.random.seed(13)
        want = getattr(pyref, name)(*args)
        np.random.seed(13)
        got = to_host(getattr(S, name)(*args))
        if tol == 0.:
            assert_bit_equal(got, want, what=name)
            continue
        for k in range(10):
            scale = max(1., float(np.max(np.abs(want[k]))))
            t = 4e-16 * scale if k < 4 else tol
            assert np.max(np.abs(got[k] - want[k])) <= t, (name, k, np.max(np.abs(got[k] - want[k])))
    # explicit draws == the global stream
    np.random.seed(13)
    d = [np.random.rand(4001) for _ in range(3)]
    a = to_host(S.convergingbeam(8400., 200., 230., -.1, .3, 4001, 1.5, draws=d))
    np.random.seed(13)
    b = to_host(S.convergingbeam(8400., 200., 230., -.1, .3, 4001, 1.5))
    assert_bit_equal(a, b, what="draws")


def test_beam_sources_philox(pxf):
    """Device draws: shards concatenate to the single-GPU stream, and the distributions are the reference's."""
    S = pxf.sources
    N = 200_000
    for name, args in (("rectbeam", (12., 7.)), ("gaussianBeam", (.01,)),
                       ("convergingbeam", (8400., 200., 230., -.1, .3)), ("convergingbeam2", (8400., -20., 30., 190., 240.))):
        tail = (1.5,) if name.startswith("converging") else ()
        whole = to_host(getattr(S, name)(*args, N, *tail, rng="philox", seed=9))
        a = to_host(getattr(S, name)(*args, 70_001, *tail, rng="philox", seed=9, first=0))
        b = to_host(getattr(S, name)(*args, N - 70_001, *tail, rng="philox", seed=9, first=70_001))
        assert_bit_equal([np.concatenate([p, q]) for p, q in zip(a, b)], whole, what=name + " philox shards")
        other = to_host(getattr(S, name)(*args, N, *tail, rng="philox", seed=10))
        assert not np.array_equal(other[1] + other[4], whole[1] + whole[4])
        opd, x, y, z, l, m, n, ux, uy, uz = whole
        assert np.all(np.abs(l ** 2 + m ** 2 + n ** 2 - 1.) < 1e-12)
        if name == "rectbeam":
            assert x.min() >= -12. and x.max() <= 12. and abs(x.mean()) < .1 and abs(x.std() - 24. / np.sqrt(12.)) < .05
            assert y.min() >= -7. and y.max() <= 7. and abs(np.corrcoef(x, y)[0, 1]) < .01
        elif name == "gaussianBeam":
            sig = np.sin(.01) / np.sqrt(2.)
            assert abs(l.std() / sig - 1.) < .01 and abs(m.std() / sig - 1.) < .01 and abs(l.mean()) < 1e-4
            assert abs(np.corrcoef(l, m)[0, 1]) < .01
            assert abs(np.mean(np.abs(l) < sig) - .6827) < .005            # a normal, not just the right variance
        elif name == "convergingbeam":
            r = np.hypot(x, y)
            assert r.min() >= 200. - 1e-9 and r.max() <= 230. + 1e-9 and np.all(z == 8400.)
            u = (r ** 2 - 200. ** 2) / (230. ** 2 - 200. ** 2)
            th = np.arctan2(y, x)
            assert abs(u.mean() - .5) < .005 and th.min() >= -.1 - 1e-12 and th.max() <= .3 + 1e-12
            # the rays converge to the origin: footprint at z = 0 is set by the Lorentzian scatter (median |.| = lscat)
            x0 = x - l / n * z
            assert np.median(np.abs(np.hypot(x0, y - m / n * z))) < .2
        else:
            assert x.min() >= -20. and x.max() <= 30. and y.min() >= 190. and y.max() <= 240.


def test_pointto_applyt_indangle_kernels(pxf):
    """transformations.pointTo (:91-100) bit for bit; applyT (:257-280) against numpy's matrix product;
    analyses.indAngle (:164-182) with mask / index selections."""
    A, T = pxf.analyses, pxf.transformations
    np.random.seed(21)
    rays = pyref.subannulus(200., 230., .4, 5003, -1.)
    pyref.transform(rays, 1., -2., 30., .01, -.02, .3)
    rays[7][:] = np.sin(.1) * np.cos(np.linspace(0, 6, 5003))
    rays[8][:] = np.sin(.1) * np.sin(np.linspace(0, 6, 5003))
    rays[9][:] = np.cos(.1)
    for rev in (-1., 1.):
        want = copy(rays)
        R = np.sqrt((want[1] - 3.) ** 2 + (want[2] + 4.) ** 2 + (want[3] - 8000.) ** 2)
        want[4] = rev * (want[1] - 3.) / R
        want[5] = rev * (want[2] + 4.) / R
        want[6] = rev * (want[3] - 8000.) / R
        dev = to_dev(rays)
        T.pointTo(dev, 3., -4., 8000., reverse=rev)
        assert_bit_equal(to_host(dev), want, what="pointTo")
    coords = T.newCoords()
    T._update_coords_fwd(coords, 1., 2., 3., .1, .2, .3)
    T._update_coords_fwd(coords, -5., 0., 100., -.3, .05, 1.)
    for inv in (False, True):
        i = 2 if inv else 0
        on = np.ones(5003)
        pos = np.dot(coords[i + 1], [rays[1], rays[2], rays[3], on])[:3]
        wav = np.dot(coords[i], [rays[4], rays[5], rays[6], on])[:3]
        nrm = np.dot(coords[i], [rays[7], rays[8], rays[9], on])[:3]
        dev = to_dev(rays)
        got = to_host(T.applyT(dev, coords, inverse=inv))
        assert_bit_equal(to_host(dev), rays, what="applyT leaves its input alone")
        assert np.array_equal(got[0], rays[0])
        for k in range(3):
            assert np.max(np.abs(got[1 + k] - pos[k])) <= 1e-12 * max(1., np.max(np.abs(pos[k])))
            assert np.max(np.abs(got[4 + k] - wav[k])) <= 1e-15 * 4
            assert np.max(np.abs(got[7 + k] - nrm[k])) <= 1e-15 * 4
    dev = to_dev(rays)
    ia = np.arccos(rays[4] * rays[7] + rays[5] * rays[8] + rays[6] * rays[9])
    mask = rays[1] > 215.
    idx = np.where(mask)[0][::-3]
    assert np.allclose(A.indAngle(dev, ind=mask).cpu().numpy(), ia[mask], rtol=0, atol=1e-14)
    assert np.allclose(A.indAngle(dev, ind=idx).cpu().numpy(), ia[idx], rtol=0, atol=1e-14)
    assert np.allclose(A.indAngle(dev, ind=np.where(mask)).cpu().numpy(), ia[mask], rtol=0, atol=1e-14)
    got = A.indAngle(dev, ind=mask, normal=(0., .6, .8)).cpu().numpy()
    assert np.allclose(got, np.arccos(.6 * rays[5][mask] + .8 * rays[6][mask]), rtol=0, atol=1e-14)
    assert A.indAngle(dev, ind=np.zeros(5003, dtype=bool)).shape[0] == 0
    # measureOPD (analyses.py:232-244): an (x,y,z) triple or a ten-row ray
    want = np.sqrt((rays[1] - 3.) ** 2 + (rays[2] + 4.) ** 2 + (rays[3] - 8000.) ** 2)
    assert np.array_equal(A.measureOPD(dev, (3., -4., 8000.)).cpu().numpy(), want)
    assert np.array_equal(A.measureOPD(dev, [0., 3., -4., 8000., 0., 0., 1., 0., 0., 1.]).cpu().numpy(), want)


@pytest.mark.parametrize("n", [1, 2, 31, 3072, 3073, 100_000, 2_000_003])
def test_argsort_sizes_and_degenerate_keys(pxf, n):
    """The one-sweep radix sort (analyses.py:76 argsort): sizes around the tile, all-equal keys (no pass runs), keys
    that differ in one byte only, heavy ties (stability), negative / signed-zero / inf / NaN keys."""
    import torch
    rng = np.random.default_rng(n)
    cases = {
        "uniform": rng.uniform(0, 1e-3, n),
        "equal": np.full(n, 3.25),
        "one_byte": np.float64(1.) + rng.integers(0, 256, n) * 2. ** -52,
        "ties": rng.integers(-3, 4, n).astype(np.float64),
        "mixed": rng.normal(0, 1, n) * 10. ** rng.integers(-300, 300, n),
    }
    cases["mixed"][::7] = -0.
    cases["mixed"][3::11] = 0.
    cases["mixed"][5::13] = np.inf
    cases["mixed"][6::17] = -np.inf
    cases["mixed"][1::19] = np.nan
    for name, kc in cases.items():
        ks, idx = pxf.analyses.argsort(torch.from_numpy(kc).cuda())
        kn, ii = ks.cpu().numpy(), idx.cpu().numpy()
        want = np.argsort(kc, kind="stable")
        assert np.array_equal(ii, want), name
        assert np.array_equal(kn.view(np.int64)[~np.isnan(kn)], kc[want].view(np.int64)[~np.isnan(kc[want])]), name
        assert np.isnan(kn).sum() == np.isnan(kc).sum(), name


def test_argsort_at_bench_scale(pxf):
    """1e8 keys (the radii of a config-5 shard): size-independent properties of the one-sweep sort -- sorted, a permutation,
    keys[idx] == sorted keys bit for bit, equal keys in input order."""
    import torch
    n = 100_000_000
    g = torch.Generator(device="cuda").manual_seed(3)
    keys = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    keys[::1000] = keys[0]                                   # 1e5 exact ties
    keys.mul_(1e-3)
    ks, idx = pxf.analyses.argsort(keys)
    assert bool((ks[1:] >= ks[:-1]).all())
    assert bool((keys[idx] == ks).all())
    seen = torch.zeros(n, dtype=torch.bool, device="cuda")
    seen[idx] = True
    assert bool(seen.all())
    tie = ks[1:] == ks[:-1]
    assert int(tie.sum()) >= 99_000 and bool((idx[1:][tie] > idx[:-1][tie]).all())
