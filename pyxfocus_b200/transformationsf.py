"""Replacement for the reference's f2py module ``transformationsf`` (built from
``transformationsf.f95`` by ``compiletrace.sh:2``).

Same routine names (lower-cased), same positional argument order, in-place mutation,
``None`` return -- but the arrays are rows of the device ray bundle and every routine is one
CUDA kernel launch behind the C ABI in ``include/pxf.h``.  Extension over f2py: an optional
``mask=`` (uint8/bool per ray) runs the routine only where the mask is set, which is what the
reference does with gather -> Fortran -> scatter (``transformations.py:20-27``).
"""
from . import _lib
from ._call import Staged, run


class error(_lib.PxfError):
    """Name kept for code that catches ``transformationsf.error``."""


def _chk(st, num):
    if num is not None and int(num) != st.num:
        raise ValueError("shape(x,0)==num failed")


def transform(x, y, z, l, m, n, ux, uy, uz, tx, ty, tz, rx, ry, rz, num=None, mask=None):
    """transformationsf.f95:134-163"""
    st = Staged()
    p = [st.inout(a) for a in (x, y, z, l, m, n, ux, uy, uz)]
    _chk(st, num)
    run(_lib.lib().pxf_transform, st, *p, st.num, tx, ty, tz, rx, ry, rz, st.mask(mask), st.stream())


def itransform(x, y, z, l, m, n, ux, uy, uz, tx, ty, tz, rx, ry, rz, num=None, mask=None):
    """transformationsf.f95:168-201"""
    st = Staged()
    p = [st.inout(a) for a in (x, y, z, l, m, n, ux, uy, uz)]
    _chk(st, num)
    run(_lib.lib().pxf_itransform, st, *p, st.num, tx, ty, tz, rx, ry, rz, st.mask(mask), st.stream())


def reflect(l, m, n, ux, uy, uz, num=None, mask=None):
    """transformationsf.f95:60-79"""
    st = Staged()
    p = [st.inout(a) for a in (l, m, n, ux, uy, uz)]
    _chk(st, num)
    run(_lib.lib().pxf_reflect, st, *p, st.num, st.mask(mask), st.stream())


def refract(l, m, n, ux, uy, uz, n1, n2, num=None, mask=None):
    """transformationsf.f95:82-130"""
    st = Staged()
    p = [st.inout(a) for a in (l, m, n, ux, uy, uz)]
    _chk(st, num)
    run(_lib.lib().pxf_refract, st, *p, st.num, n1, n2, st.mask(mask), st.stream())


def radgrat(x, y, l, m, n, wave, dpermm, order, num=None, mask=None):
    """transformationsf.f95:205-238 (scalar wavelength)"""
    st = Staged()
    lmn = [st.inout(a) for a in (l, m, n)]
    xy = [st.input(a) for a in (x, y)]
    _chk(st, num)
    run(_lib.lib().pxf_radgrat, st, *xy, *lmn, float(wave), st.num, dpermm, order, st.mask(mask), st.stream())


def radgratw(x, y, l, m, n, wave, dpermm, order, num=None, mask=None):
    """transformationsf.f95:242-272 (per-ray wavelength, sign of n taken from y)"""
    st = Staged()
    lmn = [st.inout(a) for a in (l, m, n)]
    xy = [st.input(a) for a in (x, y)]
    w = st.input(wave)
    _chk(st, num)
    run(_lib.lib().pxf_radgratw, st, *xy, *lmn, w, st.num, dpermm, order, st.mask(mask), st.stream())


def grat(x, y, l, m, n, d, order, wave, num=None, mask=None):
    """transformationsf.f95:277-305"""
    st = Staged()
    lmn = [st.inout(a) for a in (l, m, n)]
    xy = [st.input(a) for a in (x, y)]
    o = st.input(order)
    w = st.input(wave)
    _chk(st, num)
    run(_lib.lib().pxf_grat, st, *xy, *lmn, st.num, d, o, w, st.mask(mask), st.stream())
