"""Shared helpers for the parity tests (CUDA engine vs CPU oracle)."""
import numpy as np

from oracle import chains, pyref  # noqa: F401
from oracle import f2py as of

ROWS = ["opd", "x", "y", "z", "l", "m", "n", "ux", "uy", "uz"]


def copy(rays):
    return [np.array(r, dtype=np.float64, copy=True) for r in rays]


def rows_of(mat):
    """[10,N] golden matrix -> list of ten contiguous rows."""
    return [np.ascontiguousarray(mat[i]) for i in range(10)]


def random_bundle(num, seed=0):
    """Generic well-conditioned bundle: positions O(100), unit directions, unit normals."""
    rng = np.random.default_rng(seed)
    pos = rng.normal(0., 100., (3, num))
    d = rng.normal(0., 1., (3, num))
    d /= np.sqrt((d ** 2).sum(0))
    u = rng.normal(0., 1., (3, num))
    u /= np.sqrt((u ** 2).sum(0))
    opd = rng.normal(0., 1., num)
    return [np.ascontiguousarray(a) for a in (opd, *pos, *d, *u)]


def assert_bit_equal(got, want, rows=range(10), what=""):
    for k in rows:
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, "%s row %s: shape %s vs %s" % (what, ROWS[k], a.shape, b.shape)
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        if not same.all():
            i = int(np.argmin(same))
            raise AssertionError("%s row %s: %d of %d entries differ, first at %d: %r vs %r"
                                 % (what, ROWS[k], (~same).sum(), a.size, i, a[i], b[i]))


def assert_close(got, want, pos_scale, tol=1e-12, rows=range(10), what=""):
    """Positions/opd: |delta| <= tol * pos_scale (the system's length scale: every position on
    the path is computed from coordinates of that magnitude, so this is 'relative error' in
    the only sense that survives cancellation at a focus).  Direction cosines and normals are
    components of unit vectors: absolute error <= tol.  NaNs must coincide."""
    for k in rows:
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, "%s row %s: shape" % (what, ROWS[k])
        na, nb = np.isnan(a), np.isnan(b)
        assert np.array_equal(na, nb), "%s row %s: NaN pattern differs (%d vs %d)" % (what, ROWS[k], na.sum(), nb.sum())
        scale = pos_scale if k < 4 else 1.0
        err = np.abs(np.where(na, 0., a - b))
        worst = err.max() if err.size else 0.
        assert worst <= tol * scale, "%s row %s: max |delta| %.3e > %.1e * %.3g" % (what, ROWS[k], worst, tol, scale)


def steps_to_program(steps):
    from pyxfocus_b200 import Program
    p = Program()
    for name, a in steps:
        getattr(p, name)(*a)
    return p


def run_steps_gpu(rays, steps):
    """Per-routine execution through the f2py-shaped modules (one kernel per routine)."""
    from pyxfocus_b200 import surfacesf as S, transformationsf as T, woltsurf as W
    from pyxfocus_b200 import Program
    opd, x, y, z, l, m, n, ux, uy, uz = rays
    for name, a in steps:
        if name in ("transform", "itransform"):
            getattr(T, name)(x, y, z, l, m, n, ux, uy, uz, *a)
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
        elif name in ("conicopd", "wolterprimaryopd"):
            getattr(S if name == "conicopd" else W, name)(opd, x, y, z, l, m, n, ux, uy, uz, *a)
        elif name in ("wolterprimary", "woltersecondary", "woltersine", "wsprimary", "wssecondary", "spocone"):
            getattr(W, name)(x, y, z, l, m, n, ux, uy, uz, *a)
        elif name == "kick":
            Program().kick(*a).run(rays)
        else:
            raise ValueError(name)


def to_dev(rays, device="cuda"):
    from pyxfocus_b200 import sources
    return sources.from_numpy(rays, device=device)


def to_host(rays):
    from pyxfocus_b200 import sources
    return sources.to_numpy(rays)


class UploadedSources:
    """``sources`` stand-in for bit-level chain tests: the ORACLE's numpy source formulas (oracle/pyref.py, i.e. the
    reference's: numpy cos/sin), uploaded.  The product's own numpy-seeded sources evaluate cos/sin on the device
    and agree to 1e-12 (test_sources_from_numpy_seeds); with identical input rays every algebraic chain must then
    agree bit for bit."""

    def __init__(self, device="cuda"):
        self.device = device

    def _up(self, rays, out):
        import torch
        if out is None:
            return to_dev(rays, self.device)
        for k in range(10):
            out[k].copy_(torch.from_numpy(np.ascontiguousarray(rays[k])))
        return out

    def subannulus(self, rin, rout, dphi, num, zhat=1., device=None, out=None, **kw):
        return self._up(pyref.subannulus(rin, rout, dphi, int(num), zhat=zhat), out)

    def annulus(self, rin, rout, num, zhat=-1., device=None, out=None, **kw):
        return self._up(pyref.annulus(rin, rout, int(num), zhat=zhat), out)

    def pointsource(self, ang, num, device=None, out=None, **kw):
        return self._up(pyref.pointsource(ang, int(num)), out)


def exact_product_api(ex, device=None):
    """The product's modules with the sources swapped for ``UploadedSources``."""
    from types import SimpleNamespace
    import pyxfocus_b200 as pxf
    mods = SimpleNamespace(sources=UploadedSources(), transformations=pxf.transformations, surfaces=pxf.surfaces,
                           analyses=pxf.analyses, conicsolve=pxf.conicsolve)
    return ex.make_api(mods, ex.TorchXP(device), "pyxfocus_b200 (oracle sources)")
