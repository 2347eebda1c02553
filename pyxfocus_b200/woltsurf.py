"""Replacement for the reference's f2py module ``woltsurf`` (``woltsurf.f95``,
``compiletrace.sh:4``).  See ``transformationsf`` for the conventions."""
import numpy as np
import torch

from . import _lib
from ._call import Staged, run


class error(_lib.PxfError):
    pass


def _nine(fn_name, arrs, scalars, num, mask):
    st = Staged()
    p = [st.inout(a) for a in arrs]
    if num is not None and int(num) != st.num:
        raise ValueError("shape(x,0)==num failed")
    run(getattr(_lib.lib(), fn_name), st, *p, st.num, *scalars, st.mask(mask), st.stream())


def wolterprimary(x, y, z, l, m, n, ux, uy, uz, r0, z0, psi, num=None, mask=None):
    """woltsurf.f95:7-54"""
    _nine("pxf_wolterprimary", (x, y, z, l, m, n, ux, uy, uz), (r0, z0, psi), num, mask)


def wolterprimaryopd(opd, x, y, z, l, m, n, ux, uy, uz, r0, z0, psi, nr, num=None, mask=None):
    """woltsurf.f95:60-108"""
    _nine("pxf_wolterprimaryopd", (opd, x, y, z, l, m, n, ux, uy, uz), (r0, z0, psi, nr), num, mask)


def woltersecondary(x, y, z, l, m, n, ux, uy, uz, r0, z0, psi, num=None, mask=None):
    """woltsurf.f95:114-161"""
    _nine("pxf_woltersecondary", (x, y, z, l, m, n, ux, uy, uz), (r0, z0, psi), num, mask)


def woltersine(x, y, z, l, m, n, ux, uy, uz, r0, z0, amp, freq, num=None, mask=None):
    """woltsurf.f95:167-215"""
    _nine("pxf_woltersine", (x, y, z, l, m, n, ux, uy, uz), (r0, z0, amp, freq), num, mask)


def wsprimary(x, y, z, l, m, n, ux, uy, uz, alpha, z0, psi, num=None, mask=None):
    """woltsurf.f95:387-476"""
    _nine("pxf_wsprimary", (x, y, z, l, m, n, ux, uy, uz), (alpha, z0, psi), num, mask)


def wssecondary(x, y, z, l, m, n, ux, uy, uz, alpha, z0, psi, num=None, mask=None):
    """woltsurf.f95:484-588"""
    _nine("pxf_wssecondary", (x, y, z, l, m, n, ux, uy, uz), (alpha, z0, psi), num, mask)


def spocone(x, y, z, l, m, n, ux, uy, uz, r0, tg, num=None, mask=None):
    """woltsurf.f95:591-638"""
    _nine("pxf_spocone", (x, y, z, l, m, n, ux, uy, uz), (r0, tg), num, mask)


def wsprimaryback(x, y, z, l, m, n, ux, uy, uz, alpha, z0, psi, thick, num=None, mask=None):
    """woltsurf.f95:726-815"""
    _nine("pxf_wsprimaryback", (x, y, z, l, m, n, ux, uy, uz), (alpha, z0, psi, thick), num, mask)


def wssecondaryback(x, y, z, l, m, n, ux, uy, uz, alpha, z0, psi, thick, num=None, mask=None):
    """woltsurf.f95:824-933"""
    _nine("pxf_wssecondaryback", (x, y, z, l, m, n, ux, uy, uz), (alpha, z0, psi, thick), num, mask)


def _ll(fn_name, arrs, scalars, coeff, axial, az, num, cnum, mask):
    def host(a, dtype):
        if isinstance(a, torch.Tensor):
            a = a.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(a).ravel(), dtype=dtype)
    c, ax, azz = host(coeff, np.float64), host(axial, np.int32), host(az, np.int32)
    if cnum is not None and int(cnum) != c.shape[0]:
        raise ValueError("shape(coeff,0)==cnum failed")
    if ax.shape[0] != c.shape[0] or azz.shape[0] != c.shape[0]:
        raise ValueError("shape(axial,0)==cnum failed")
    st = Staged()
    p = [st.inout(a) for a in arrs]
    if num is not None and int(num) != st.num:
        raise ValueError("shape(x,0)==num failed")
    run(getattr(_lib.lib(), fn_name), st, *p, st.num, *scalars, c.ctypes.data, ax.ctypes.data, azz.ctypes.data,
        c.shape[0], st.mask(mask), st.stream())


def wolterprimll(x, y, z, l, m, n, ux, uy, uz, r0, z0, zmax, zmin, dphi, coeff, axial, az, num=None, cnum=None,
                 mask=None):
    """woltsurf.f95:219-288"""
    _ll("pxf_wolterprimll", (x, y, z, l, m, n, ux, uy, uz), (r0, z0, zmax, zmin, dphi), coeff, axial, az, num, cnum, mask)


def woltersecll(x, y, z, l, m, n, ux, uy, uz, r0, z0, psi, zmax, zmin, dphi, coeff, axial, az, num=None, cnum=None,
                mask=None):
    """woltsurf.f95:293-379"""
    _ll("pxf_woltersecll", (x, y, z, l, m, n, ux, uy, uz), (r0, z0, psi, zmax, zmin, dphi), coeff, axial, az, num, cnum,
        mask)


def ellipsoidwoltll(x, y, z, l, m, n, ux, uy, uz, r0, z0, psi, s, zmax, zmin, dphi, coeff, axial, az, num=None,
                    cnum=None, mask=None):
    """woltsurf.f95:643-718"""
    _ll("pxf_ellipsoidwoltll", (x, y, z, l, m, n, ux, uy, uz), (r0, z0, psi, s, zmax, zmin, dphi), coeff, axial, az, num,
        cnum, mask)
