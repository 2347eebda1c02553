"""The BASELINE configurations as op chains over the CPU oracle (TEST INFRASTRUCTURE ONLY).

Each chain is written twice: ``*_cpu`` drives the oracle's f2py-shaped modules on numpy
arrays exactly as the reference's example scripts drive the Fortran (call sites cited), and
``*_program`` returns the same chain as a list of (routine name, args) tuples that the tests
and bench turn into a ``pyxfocus_b200.Program`` or into per-routine calls.
"""
import numpy as np

from . import f2py as _f
from . import pyref

T, S, W, Z = _f.transformationsf, _f.surfacesf, _f.woltsurf, _f.zernsurf


# ------------------------------------------------------------------ config 1: Wolter-I pair
# examples/axro/singlePassAlignment.py:246-269 with secalign=0
WOLTER1 = dict(r0=220., z0=8400., psi=1., rin=220., rout=220.6)


def wolter1_source(num, seed=0, dphi=2 * np.pi):
    np.random.seed(seed)
    return pyref.subannulus(WOLTER1["rin"], WOLTER1["rout"], dphi, num, zhat=-1.)


def wolter1_steps(r0=220., z0=8400., psi=1.):
    """(routine, f2py-level scalar args) in call order."""
    return [("transform", (0., 0., z0, 0., 0., 0.)),      # tran.transform(rays,0,0,-8400,0,0,0) negated
            ("wolterprimary", (r0, z0, psi)),
            ("reflect", ()),
            ("woltersecondary", (r0, z0, psi)),
            ("reflect", ()),
            ("flat", ())]


def run_steps_cpu(rays, steps):
    """Execute a step list on the oracle (numpy arrays, in place)."""
    opd, x, y, z, l, m, n, ux, uy, uz = rays
    for name, a in steps:
        if name == "transform":
            T.transform(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "itransform":
            T.itransform(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "reflect":
            T.reflect(l, m, n, ux, uy, uz)
        elif name == "refract":
            T.refract(l, m, n, ux, uy, uz, *a)
        elif name == "radgrat":
            T.radgrat(x, y, l, m, n, *a)
        elif name == "flat":
            S.flat(x, y, z, l, m, n, ux, uy, uz)
        elif name == "flatopd":
            S.flatopd(x, y, z, l, m, n, ux, uy, uz, opd, *a)
        elif name == "conic":
            S.conic(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "conicopd":
            S.conicopd(opd, x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "wolterprimary":
            W.wolterprimary(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "wolterprimaryopd":
            W.wolterprimaryopd(opd, x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "woltersecondary":
            W.woltersecondary(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "woltersine":
            W.woltersine(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "wsprimary":
            W.wsprimary(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "wssecondary":
            W.wssecondary(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "spocone":
            W.spocone(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "kick":
            dl, dm, sn = a
            l += dl
            m += dm
            n[:] = sn * np.sqrt(1. - l ** 2 - m ** 2)
        else:
            raise ValueError(name)
    return rays


def wolter1_cpu(rays):
    """Config 1 end to end on the oracle: trace + hpd."""
    run_steps_cpu(rays, wolter1_steps())
    return pyref.hpd(rays)


# ------------------------------------------------------------------ config 2: Wolter-Schwarzschild
# examples/axro/axialHeights.py:77-113 (traceZeta), R0=220, Z0=1e4, psi=1
WS = dict(r0=220., z0=1.e4, psi=1., L=200., az=100.)


def ws_primrad(zz, r0=220., z0=1.e4, psi=1.):
    """Radius of the W-S primary at height zz: one horizontal ray traced to the surface
    (examples/axro/axialHeights.py:51-62, wsPrimrad)."""
    np.random.seed(0)
    ray = pyref.pointsource(0., 1)
    pyref.transform(ray, 0, 0, 0, 0, -np.pi / 2, 0)
    pyref.transform(ray, -r0 - 2., 0, -zz, 0, 0, 0)
    alpha = pyref.woltparam(r0, z0)[0]
    W.wsprimary(*ray[1:], alpha, z0, psi)
    return ray[1][0]


def ws_aperture(r0=220., z0=1.e4, psi=1., L=200., pmin=None):
    """Aperture radii (a0, a1) of axialHeights.py:86-87; SURVEY.md 3.2 measured
    220.13737656836065 and 221.23348169132342 for the default shell."""
    pmin = z0 + 25. if pmin is None else pmin
    return ws_primrad(pmin, r0, z0, psi), ws_primrad(pmin + L, r0, z0, psi)


def ws_source(num, seed=0, r0=220., z0=1.e4, psi=1., L=200., az=100.):
    a0, a1 = ws_aperture(r0, z0, psi, L)
    np.random.seed(seed)
    return pyref.subannulus(a0, a1, az / r0, num)


def ws_steps(theta, r0=220., z0=1.e4, psi=1.):
    """transform -> wsPrimary -> field kick -> reflect -> wsSecondary -> reflect."""
    alpha = pyref.woltparam(r0, z0)[0]
    return [("transform", (0., 0., z0, 0., 0., 0.)),
            ("wsprimary", (alpha, z0, psi)),
            ("kick", (np.sin(theta), 0., -1.)),
            ("reflect", ()),
            ("wssecondary", (alpha, z0, psi)),
            ("reflect", ())]


# ------------------------------------------------------------------ Zernike tables (config 3)
def zernike_orders(nmax=7):
    """Explicit (rorder, aorder) for all terms up to radial order nmax: n ascending, |m|
    ascending with the cosine (+m) term before the sine (-m) term.  36 terms for nmax=7.
    (The reference's default ordering lives in an un-vendored module; tests pass orders
    explicitly)"""
    ro, ao = [], []
    for n in range(nmax + 1):
        for m in range(n % 2, n + 1, 2):
            if m == 0:
                ro.append(n); ao.append(0)
            else:
                ro.append(n); ao.append(m)
                ro.append(n); ao.append(-m)
    return np.array(ro, dtype=np.int64), np.array(ao, dtype=np.int64)


def zernike_coeff(nterms=36, seed=0, sigma=1.e-4):
    rng = np.random.default_rng(seed)
    c = rng.normal(0., sigma, nterms)
    c[:3] = 0.          # piston / tilts zeroed
    return c
