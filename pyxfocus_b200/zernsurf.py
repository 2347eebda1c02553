"""Replacement for the reference's f2py module ``zernsurf`` (``zernsurf.f95``,
``compiletrace.sh:1``): tracezern / tracezernopd / zernphase / tracezernrot.  ``coeff`` is float64 ``intent(in)``,
``rorder``/``aorder`` are cast to int32 like f2py does (``surfaces.py:39-43`` passes int64)."""
import numpy as np
import torch

from . import _lib
from ._call import Staged, run


class error(_lib.PxfError):
    pass


def _host(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.ndim != 1:
        raise ValueError("expected a rank-1 array")
    return a


def _tables(coeff, rorder, aorder, arrsize):
    c = _host(coeff, np.float64)
    r = _host(rorder, np.int32)
    a = _host(aorder, np.int32)
    if arrsize is not None and int(arrsize) != c.shape[0]:
        raise ValueError("shape(coeff,0)==arrsize failed")
    if r.shape[0] != c.shape[0] or a.shape[0] != c.shape[0]:
        raise ValueError("shape(rorder,0)==arrsize failed")
    return c, r, a


def tracezern(x, y, z, l, m, n, ux, uy, uz, coeff, rorder, aorder, rad, num=None, arrsize=None, mask=None):
    """zernsurf.f95:8-101"""
    c, r, a = _tables(coeff, rorder, aorder, arrsize)
    st = Staged()
    p = [st.inout(t) for t in (x, y, z, l, m, n, ux, uy, uz)]
    if num is not None and int(num) != st.num:
        raise ValueError("shape(x,0)==num failed")
    run(_lib.lib().pxf_tracezern, st, *p, st.num, c.ctypes.data, r.ctypes.data, a.ctypes.data, c.shape[0],
        rad, st.mask(mask), st.stream())


def tracezernopd(opd, x, y, z, l, m, n, ux, uy, uz, coeff, rorder, aorder, rad, nr, num=None, arrsize=None,
                 mask=None):
    """zernsurf.f95:108-203"""
    c, r, a = _tables(coeff, rorder, aorder, arrsize)
    st = Staged()
    p = [st.inout(t) for t in (opd, x, y, z, l, m, n, ux, uy, uz)]
    if num is not None and int(num) != st.num:
        raise ValueError("shape(x,0)==num failed")
    run(_lib.lib().pxf_tracezernopd, st, *p, st.num, c.ctypes.data, r.ctypes.data, a.ctypes.data, c.shape[0],
        rad, nr, st.mask(mask), st.stream())


def zernphase(opd, x, y, z, l, m, n, ux, uy, uz, coeff, rorder, aorder, rad, wave, num=None, arrsize=None, mask=None):
    """zernsurf.f95:206-250"""
    c, r, a = _tables(coeff, rorder, aorder, arrsize)
    st = Staged()
    p = [st.inout(t) for t in (opd, x, y, z, l, m, n, ux, uy, uz)]
    if num is not None and int(num) != st.num:
        raise ValueError("shape(x,0)==num failed")
    run(_lib.lib().pxf_zernphase, st, *p, st.num, c.ctypes.data, r.ctypes.data, a.ctypes.data, c.shape[0],
        rad, wave, st.mask(mask), st.stream())


def tracezernrot(x, y, z, l, m, n, ux, uy, uz, coeff1, rorder1, aorder1, coeff2, rorder2, aorder2, rad, rot,
                 num=None, arrsize1=None, arrsize2=None, mask=None):
    """zernsurf.f95:257-359"""
    c1, r1, a1 = _tables(coeff1, rorder1, aorder1, arrsize1)
    c2, r2, a2 = _tables(coeff2, rorder2, aorder2, arrsize2)
    st = Staged()
    p = [st.inout(t) for t in (x, y, z, l, m, n, ux, uy, uz)]
    if num is not None and int(num) != st.num:
        raise ValueError("shape(x,0)==num failed")
    run(_lib.lib().pxf_tracezernrot, st, *p, st.num, c1.ctypes.data, r1.ctypes.data, a1.ctypes.data, c1.shape[0],
        c2.ctypes.data, r2.ctypes.data, a2.ctypes.data, c2.shape[0], rad, rot, st.mask(mask), st.stream())
