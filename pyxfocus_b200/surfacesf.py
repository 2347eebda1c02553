"""Replacement for the reference's f2py module ``surfacesf`` (``surfacesf.f95``,
``compiletrace.sh:3``): flat / flatopd / conic / conicopd.  See ``transformationsf`` for
the conventions."""
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
