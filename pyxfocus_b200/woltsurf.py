"""Replacement for the reference's f2py module ``woltsurf`` (``woltsurf.f95``,
``compiletrace.sh:4``).  See ``transformationsf`` for the conventions."""
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
