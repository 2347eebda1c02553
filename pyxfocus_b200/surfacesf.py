"""Replacement for the reference's f2py module ``surfacesf`` (``surfacesf.f95``,
``compiletrace.sh:3``): every subroutine of the file.  See ``transformationsf`` for the
conventions."""
import numpy as np
import torch

from . import _lib
from ._call import Staged, run


class error(_lib.PxfError):
    pass


def _chk(st, num):
    if num is not None and int(num) != st.num:
        raise ValueError("shape(x,0)==num failed")


def flat(x, y, z, l, m, n, ux, uy, uz, num=None, mask=None):
    """surfacesf.f95:4-29 (the propagation distance is REAL*4 in the reference)"""
    st = Staged()
    xyz = [st.inout(a) for a in (x, y, z)]
    u = [st.inout(a) for a in (ux, uy, uz)]
    lmn = [st.input(a) for a in (l, m, n)]
    _chk(st, num)
    run(_lib.lib().pxf_flat, st, *xyz, *lmn, *u, st.num, st.mask(mask), st.stream())


def flatopd(x, y, z, l, m, n, ux, uy, uz, opd, nr, num=None, mask=None):
    """surfacesf.f95:32-53"""
    st = Staged()
    xyz = [st.inout(a) for a in (x, y, z)]
    u = [st.inout(a) for a in (ux, uy, uz)]
    o = st.inout(opd)
    lmn = [st.input(a) for a in (l, m, n)]
    _chk(st, num)
    run(_lib.lib().pxf_flatopd, st, *xyz, *lmn, *u, o, st.num, nr, st.mask(mask), st.stream())


def conic(x, y, z, l, m, n, ux, uy, uz, r, k, num=None, mask=None):
    """surfacesf.f95:302-360"""
    st = Staged()
    p = [st.inout(a) for a in (x, y, z, l, m, n, ux, uy, uz)]
    _chk(st, num)
    run(_lib.lib().pxf_conic, st, *p, st.num, r, k, st.mask(mask), st.stream())


def conicopd(opd, x, y, z, l, m, n, ux, uy, uz, r, k, nr, num=None, mask=None):
    """surfacesf.f95:366-420"""
    st = Staged()
    p = [st.inout(a) for a in (opd, x, y, z, l, m, n, ux, uy, uz)]
    _chk(st, num)
    run(_lib.lib().pxf_conicopd, st, *p, st.num, r, k, nr, st.mask(mask), st.stream())


def _nine(fn_name, arrs, scalars, num, mask):
    st = Staged()
    p = [st.inout(a) for a in arrs]
    _chk(st, num)
    run(getattr(_lib.lib(), fn_name), st, *p, st.num, *scalars, st.mask(mask), st.stream())


def _host(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a).ravel(), dtype=dtype)


def tracesphere(x, y, z, l, m, n, ux, uy, uz, rad, num=None, mask=None):
    """surfacesf.f95:57-101"""
    _nine("pxf_tracesphere", (x, y, z, l, m, n, ux, uy, uz), (rad,), num, mask)


def tracesphereopd(opd, x, y, z, l, m, n, ux, uy, uz, rad, nr, num=None, mask=None):
    """surfacesf.f95:104-149"""
    _nine("pxf_tracesphereopd", (opd, x, y, z, l, m, n, ux, uy, uz), (rad, nr), num, mask)


def tracecyl(x, y, z, l, m, n, ux, uy, uz, rad, num=None, mask=None):
    """surfacesf.f95:153-197"""
    _nine("pxf_tracecyl", (x, y, z, l, m, n, ux, uy, uz), (rad,), num, mask)


def tracecylopd(opd, x, y, z, l, m, n, ux, uy, uz, rad, nr, num=None, mask=None):
    """surfacesf.f95:201-246"""
    _nine("pxf_tracecylopd", (opd, x, y, z, l, m, n, ux, uy, uz), (rad, nr), num, mask)


def cylconic(x, y, z, l, m, n, ux, uy, uz, rad, k, num=None, mask=None):
    """surfacesf.f95:251-296"""
    _nine("pxf_cylconic", (x, y, z, l, m, n, ux, uy, uz), (rad, k), num, mask)


def paraxial(x, y, z, l, m, n, ux, uy, uz, f, num=None, mask=None):
    """surfacesf.f95:423-440"""
    _nine("pxf_paraxial", (x, y, z, l, m, n, ux, uy, uz), (f,), num, mask)


def paraxialy(x, y, z, l, m, n, ux, uy, uz, f, num=None, mask=None):
    """surfacesf.f95:443-460"""
    _nine("pxf_paraxialy", (x, y, z, l, m, n, ux, uy, uz), (f,), num, mask)


def torus(x, y, z, l, m, n, ux, uy, uz, rin, rout, num=None, mask=None):
    """surfacesf.f95:468-508"""
    _nine("pxf_torus", (x, y, z, l, m, n, ux, uy, uz), (rin, rout), num, mask)


def conicplus(x, y, z, l, m, n, ux, uy, uz, r, k, p, num=None, np=None, mask=None):
    """surfacesf.f95:514-572 (``np`` is f2py's name for the hidden length of ``p``)"""
    pp = _host(p, "float64")
    if np is not None and int(np) != pp.shape[0]:
        raise ValueError("shape(p,0)==np failed")
    _nine("pxf_conicplus", (x, y, z, l, m, n, ux, uy, uz), (r, k, pp.ctypes.data, pp.shape[0]), num, mask)


def conicplusopd(opd, x, y, z, l, m, n, ux, uy, uz, r, k, p, nr, num=None, np=None, mask=None):
    """surfacesf.f95:578-638"""
    pp = _host(p, "float64")
    if np is not None and int(np) != pp.shape[0]:
        raise ValueError("shape(p,0)==np failed")
    _nine("pxf_conicplusopd", (opd, x, y, z, l, m, n, ux, uy, uz), (r, k, pp.ctypes.data, pp.shape[0], nr), num, mask)


def legsurf(x, y, z, l, m, n, ux, uy, uz, xwidth, ywidth, order, coeff, xo, yo, nc=None, num=None, mask=None):
    """surfacesf.f95:642-668"""
    c, a, b = _host(coeff, "float64"), _host(xo, "int32"), _host(yo, "int32")
    if nc is not None and int(nc) != c.shape[0]:
        raise ValueError("shape(coeff,0)==nc failed")
    if a.shape[0] != c.shape[0] or b.shape[0] != c.shape[0]:
        raise ValueError("shape(xo,0)==nc failed")
    st = Staged()
    p = [st.inout(t) for t in (x, y, z, l, m, n, ux, uy, uz)]
    _chk(st, num)
    run(_lib.lib().pxf_legsurf, st, *p, st.num, xwidth, ywidth, order, c.ctypes.data, a.ctypes.data, b.ctypes.data,
        c.shape[0], st.mask(mask), st.stream())
