"""f2py-shaped Python bindings over ``liboracle.so`` (TEST INFRASTRUCTURE ONLY).

Each function mirrors the signature f2py generates for the corresponding
Fortran subroutine (array-length arguments dropped, names lower-cased;
call sites ``surfaces.py:21-43,110-112,224-226,242,269,342,378,413`` and
``transformations.py:24,29,58,63,108,111,120,163,170``).  ``intent(inout)``
arrays must be 1-D contiguous float64 and are mutated in place; anything else
raises ``ValueError`` like the f2py wrapper would.
"""
import ctypes
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "pxf_oracle.c")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_d = ctypes.c_double
_i64 = ctypes.c_int64
_i32 = ctypes.c_int


def build(force=False):
    """Compile the oracle with the committed Makefile if it is stale."""
    stale = force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


_SO_CR = os.path.join(_HERE, "_build", "liboracle_cr.so")
_libs = {}


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _declare(_lib)
        _libs["system"] = _lib
    return _lib


class libm:
    """``with libm("cr"):`` -- run the oracle with every libm call correctly rounded (liboracle_cr.so: binary128
    libquadmath, one rounding; see the header of pxf_oracle.c) instead of this image's glibc.  The reference's libm is
    unpinned; for the chaotic Wolter-Schwarzschild rays the last bit of sin/atan2/pow decides discrete outcomes, and the
    correctly rounded value is the one every libm approximates."""

    def __init__(self, kind):
        if kind not in ("cr", "system"):
            raise ValueError(kind)
        self.kind = kind

    def __enter__(self):
        global _lib
        lib()
        if self.kind not in _libs:
            if not os.path.exists(_SO_CR) or os.path.getmtime(_SO_CR) < os.path.getmtime(_SRC):
                subprocess.run(["make", "-C", _HERE, "-s"], check=True)
            L = ctypes.CDLL(_SO_CR)
            _declare(L)
            _libs[self.kind] = L
        self.prev = _lib
        _lib = _libs[self.kind]
        return self

    def __exit__(self, *a):
        global _lib
        _lib = self.prev
        return False


def _declare(L):
    nine = [_dp] * 9
    L.pxfo_reflect.argtypes = [_dp] * 6 + [_i64]
    L.pxfo_refract.argtypes = [_dp] * 6 + [_i64, _d, _d]
    L.pxfo_transform.argtypes = nine + [_i64] + [_d] * 6
    L.pxfo_itransform.argtypes = nine + [_i64] + [_d] * 6
    L.pxfo_radgrat.argtypes = [_dp] * 5 + [_d, _i64, _d, _d]
    L.pxfo_radgratw.argtypes = [_dp] * 6 + [_i64, _d, _d]
    L.pxfo_grat.argtypes = [_dp] * 5 + [_i64, _d, _dp, _dp]
    L.pxfo_flat.argtypes = nine + [_i64]
    L.pxfo_flatopd.argtypes = nine + [_dp, _i64, _d]
    L.pxfo_conic.argtypes = nine + [_i64, _d, _d]
    L.pxfo_conicopd.argtypes = [_dp] * 10 + [_i64, _d, _d, _d]
    L.pxfo_tracesphere.argtypes = nine + [_i64, _d]
    L.pxfo_tracesphereopd.argtypes = [_dp] * 10 + [_i64, _d, _d]
    L.pxfo_tracecyl.argtypes = nine + [_i64, _d]
    L.pxfo_tracecylopd.argtypes = [_dp] * 10 + [_i64, _d, _d]
    L.pxfo_cylconic.argtypes = nine + [_i64, _d, _d]
    L.pxfo_paraxial.argtypes = nine + [_i64, _d]
    L.pxfo_paraxialy.argtypes = nine + [_i64, _d]
    L.pxfo_torus.argtypes = nine + [_i64, _d, _d]
    L.pxfo_conicplus.argtypes = nine + [_i64, _d, _d, _dp, _i32]
    L.pxfo_conicplusopd.argtypes = [_dp] * 10 + [_i64, _d, _d, _dp, _i32, _d]
    L.pxfo_legsurf.argtypes = nine + [_d, _d, _d, _dp, _ip, _ip, _i32, _i64]
    L.pxfo_wsprimaryback.argtypes = nine + [_i64, _d, _d, _d, _d]
    L.pxfo_wssecondaryback.argtypes = nine + [_i64, _d, _d, _d, _d]
    L.pxfo_wolterprimary.argtypes = nine + [_i64, _d, _d, _d]
    L.pxfo_wolterprimaryopd.argtypes = [_dp] * 10 + [_i64, _d, _d, _d, _d]
    L.pxfo_woltersecondary.argtypes = nine + [_i64, _d, _d, _d]
    L.pxfo_woltersine.argtypes = nine + [_i64, _d, _d, _d, _d]
    L.pxfo_wsprimary.argtypes = nine + [_i64, _d, _d, _d]
    L.pxfo_wssecondary.argtypes = nine + [_i64, _d, _d, _d]
    L.pxfo_spocone.argtypes = nine + [_i64, _d, _d]
    L.pxfo_wolterprimll.argtypes = nine + [_i64, _d, _d, _d, _d, _d, _dp, _ip, _ip, _i32]
    L.pxfo_woltersecll.argtypes = nine + [_i64, _d, _d, _d, _d, _d, _d, _dp, _ip, _ip, _i32]
    L.pxfo_ellipsoidwoltll.argtypes = nine + [_i64, _d, _d, _d, _d, _d, _d, _d, _dp, _ip, _ip, _i32]
    L.pxfo_zernset.argtypes = [_d, _d, _ip, _ip, _i32, _dp, _dp, _dp]
    L.pxfo_tracezern.argtypes = nine + [_i64, _dp, _ip, _ip, _i32, _d]
    L.pxfo_tracezernopd.argtypes = [_dp] * 10 + [_i64, _dp, _ip, _ip, _i32, _d, _d]
    L.pxfo_zernphase.argtypes = [_dp] * 10 + [_i64, _dp, _ip, _ip, _i32, _d, _d]
    L.pxfo_tracezernrot.argtypes = nine + [_i64, _dp, _ip, _ip, _i32, _dp, _ip, _ip, _i32, _d, _d]
    L.pxfo_radialpoly.argtypes = [_d, _i32, _i32]
    L.pxfo_radialpoly.restype = _d
    L.pxfo_legendre.argtypes = [_d, _i32]
    L.pxfo_legendre.restype = _d
    L.pxfo_legendrep.argtypes = [_d, _i32]
    L.pxfo_legendrep.restype = _d
    L.pxfo_woltersecondary_steps.argtypes = [_d] * 9
    L.pxfo_woltersecondary_steps.restype = _i32
    L.pxfo_reconstruct.argtypes = [_dp, _dp, _i32, _i32, _d, _d, _dp, _dp, _i32]
    L.pxfo_reconstruct.restype = _i64
    L.pxfo_southwellbin.argtypes = [_dp] * 4 + [_i64, _d, _dp, _dp, _dp, _i32, _i32]


def _io(*arrs):
    """Validate intent(inout) arrays the way f2py does and return pointers."""
    n = None
    out = []
    for a in arrs:
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.ndim != 1 or not a.flags.c_contiguous:
            raise ValueError("failed in converting argument to C/Fortran array: "
                             "intent(inout) array must be a contiguous 1-D float64 ndarray")
        if n is None:
            n = a.shape[0]
        elif a.shape[0] != n:
            raise ValueError("shape mismatch against num")
        out.append(a.ctypes.data_as(_dp))
    return n, out


def _in(a, n=None):
    """intent(in) array: f2py silently copies/casts."""
    b = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and b.shape[0] != n:
        raise ValueError("shape mismatch against num")
    return b, b.ctypes.data_as(_dp)


def _orders(rorder, aorder, arrsize):
    r = np.ascontiguousarray(rorder, dtype=np.int32)
    a = np.ascontiguousarray(aorder, dtype=np.int32)
    if r.shape[0] != arrsize or a.shape[0] != arrsize:
        raise ValueError("shape mismatch against arrsize")
    return r, a


# ---------------------------------------------------------------- transformationsf
def _transform(x, y, z, l, m, n, ux, uy, uz, tx, ty, tz, rx, ry, rz):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_transform(*p, num, tx, ty, tz, rx, ry, rz)


def _itransform(x, y, z, l, m, n, ux, uy, uz, tx, ty, tz, rx, ry, rz):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_itransform(*p, num, tx, ty, tz, rx, ry, rz)


def _reflect(l, m, n, ux, uy, uz):
    num, p = _io(l, m, n, ux, uy, uz)
    lib().pxfo_reflect(*p, num)


def _refract(l, m, n, ux, uy, uz, n1, n2):
    num, p = _io(l, m, n, ux, uy, uz)
    lib().pxfo_refract(*p, num, n1, n2)


def _radgrat(x, y, l, m, n, wave, dpermm, order):
    num, p = _io(l, m, n)
    xb, xp = _in(x, num)
    yb, yp = _in(y, num)
    lib().pxfo_radgrat(xp, yp, *p, float(wave), num, dpermm, order)


def _radgratw(x, y, l, m, n, wave, dpermm, order):
    num, p = _io(l, m, n)
    xb, xp = _in(x, num)
    yb, yp = _in(y, num)
    wb, wp = _in(wave, num)
    lib().pxfo_radgratw(xp, yp, *p, wp, num, dpermm, order)


def _grat(x, y, l, m, n, d, order, wave):
    num, p = _io(l, m, n)
    xb, xp = _in(x, num)
    yb, yp = _in(y, num)
    ob, op = _in(order, num)
    wb, wp = _in(wave, num)
    lib().pxfo_grat(xp, yp, *p, num, d, op, wp)


transformationsf = SimpleNamespace(transform=_transform, itransform=_itransform, reflect=_reflect,
                                   refract=_refract, radgrat=_radgrat, radgratw=_radgratw, grat=_grat)


# ---------------------------------------------------------------- surfacesf
def _flat(x, y, z, l, m, n, ux, uy, uz):
    num, p = _io(x, y, z, ux, uy, uz)
    lb, lp = _in(l, num)
    mb, mp = _in(m, num)
    nb, np_ = _in(n, num)
    lib().pxfo_flat(p[0], p[1], p[2], lp, mp, np_, p[3], p[4], p[5], num)


def _flatopd(x, y, z, l, m, n, ux, uy, uz, opd, nr):
    num, p = _io(x, y, z, ux, uy, uz, opd)
    lb, lp = _in(l, num)
    mb, mp = _in(m, num)
    nb, np_ = _in(n, num)
    lib().pxfo_flatopd(p[0], p[1], p[2], lp, mp, np_, p[3], p[4], p[5], p[6], num, nr)


def _conic(x, y, z, l, m, n, ux, uy, uz, r, k):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_conic(*p, num, r, k)


def _conicopd(opd, x, y, z, l, m, n, ux, uy, uz, r, k, nr):
    num, p = _io(opd, x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_conicopd(*p, num, r, k, nr)


def _nine_scalars(name):
    def f(x, y, z, l, m, n, ux, uy, uz, *scalars):
        num, p = _io(x, y, z, l, m, n, ux, uy, uz)
        getattr(lib(), name)(*p, num, *scalars)
    return f


def _ten_scalars(name):
    def f(opd, x, y, z, l, m, n, ux, uy, uz, *scalars):
        num, p = _io(opd, x, y, z, l, m, n, ux, uy, uz)
        getattr(lib(), name)(*p, num, *scalars)
    return f


def _conicplus(x, y, z, l, m, n, ux, uy, uz, r, k, p):
    num, q = _io(x, y, z, l, m, n, ux, uy, uz)
    pb, pp = _in(p)
    lib().pxfo_conicplus(*q, num, r, k, pp, pb.shape[0])


def _conicplusopd(opd, x, y, z, l, m, n, ux, uy, uz, r, k, p, nr):
    num, q = _io(opd, x, y, z, l, m, n, ux, uy, uz)
    pb, pp = _in(p)
    lib().pxfo_conicplusopd(*q, num, r, k, pp, pb.shape[0], nr)


def _legsurf(x, y, z, l, m, n, ux, uy, uz, xwidth, ywidth, order, coeff, xo, yo):
    num, q = _io(x, y, z, l, m, n, ux, uy, uz)
    cb, cp = _in(coeff)
    a, b = _orders(xo, yo, cb.shape[0])
    lib().pxfo_legsurf(*q, xwidth, ywidth, order, cp, a.ctypes.data_as(_ip), b.ctypes.data_as(_ip), cb.shape[0], num)


surfacesf = SimpleNamespace(flat=_flat, flatopd=_flatopd, conic=_conic, conicopd=_conicopd,
                            tracesphere=_nine_scalars("pxfo_tracesphere"),
                            tracesphereopd=_ten_scalars("pxfo_tracesphereopd"),
                            tracecyl=_nine_scalars("pxfo_tracecyl"), tracecylopd=_ten_scalars("pxfo_tracecylopd"),
                            cylconic=_nine_scalars("pxfo_cylconic"), paraxial=_nine_scalars("pxfo_paraxial"),
                            paraxialy=_nine_scalars("pxfo_paraxialy"), torus=_nine_scalars("pxfo_torus"),
                            conicplus=_conicplus, conicplusopd=_conicplusopd, legsurf=_legsurf)


# ---------------------------------------------------------------- woltsurf
def _wolterprimary(x, y, z, l, m, n, ux, uy, uz, r0, z0, psi):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_wolterprimary(*p, num, r0, z0, psi)


def _wolterprimaryopd(opd, x, y, z, l, m, n, ux, uy, uz, r0, z0, psi, nr):
    num, p = _io(opd, x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_wolterprimaryopd(*p, num, r0, z0, psi, nr)


def _woltersecondary(x, y, z, l, m, n, ux, uy, uz, r0, z0, psi):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_woltersecondary(*p, num, r0, z0, psi)


def _woltersine(x, y, z, l, m, n, ux, uy, uz, r0, z0, amp, freq):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_woltersine(*p, num, r0, z0, amp, freq)


def _wsprimary(x, y, z, l, m, n, ux, uy, uz, alpha, z0, psi):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_wsprimary(*p, num, alpha, z0, psi)


def _wssecondary(x, y, z, l, m, n, ux, uy, uz, alpha, z0, psi):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_wssecondary(*p, num, alpha, z0, psi)


def _spocone(x, y, z, l, m, n, ux, uy, uz, r0, tg):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    lib().pxfo_spocone(*p, num, r0, tg)


def _ll_tables(coeff, axial, az):
    cb, cp = _in(coeff)
    a, b = _orders(axial, az, cb.shape[0])
    return cb, cp, a, b


def _wolterprimll(x, y, z, l, m, n, ux, uy, uz, r0, z0, zmax, zmin, dphi, coeff, axial, az):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    cb, cp, a, b = _ll_tables(coeff, axial, az)
    lib().pxfo_wolterprimll(*p, num, r0, z0, zmax, zmin, dphi, cp, a.ctypes.data_as(_ip), b.ctypes.data_as(_ip),
                            cb.shape[0])


def _woltersecll(x, y, z, l, m, n, ux, uy, uz, r0, z0, psi, zmax, zmin, dphi, coeff, axial, az):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    cb, cp, a, b = _ll_tables(coeff, axial, az)
    lib().pxfo_woltersecll(*p, num, r0, z0, psi, zmax, zmin, dphi, cp, a.ctypes.data_as(_ip),
                           b.ctypes.data_as(_ip), cb.shape[0])


def _ellipsoidwoltll(x, y, z, l, m, n, ux, uy, uz, r0, z0, psi, s, zmax, zmin, dphi, coeff, axial, az):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    cb, cp, a, b = _ll_tables(coeff, axial, az)
    lib().pxfo_ellipsoidwoltll(*p, num, r0, z0, psi, s, zmax, zmin, dphi, cp, a.ctypes.data_as(_ip),
                               b.ctypes.data_as(_ip), cb.shape[0])


woltsurf = SimpleNamespace(wolterprimll=_wolterprimll, woltersecll=_woltersecll, ellipsoidwoltll=_ellipsoidwoltll,
                           wolterprimary=_wolterprimary, wolterprimaryopd=_wolterprimaryopd,
                           woltersecondary=_woltersecondary, woltersine=_woltersine,
                           wsprimary=_wsprimary, wssecondary=_wssecondary, spocone=_spocone,
                           wsprimaryback=_nine_scalars("pxfo_wsprimaryback"),
                           wssecondaryback=_nine_scalars("pxfo_wssecondaryback"))


# ---------------------------------------------------------------- zernsurf
def _tracezern(x, y, z, l, m, n, ux, uy, uz, coeff, rorder, aorder, rad):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    cb, cp = _in(coeff)
    r, a = _orders(rorder, aorder, cb.shape[0])
    lib().pxfo_tracezern(*p, num, cp, r.ctypes.data_as(_ip), a.ctypes.data_as(_ip), cb.shape[0], rad)


def _tracezernopd(opd, x, y, z, l, m, n, ux, uy, uz, coeff, rorder, aorder, rad, nr):
    num, p = _io(opd, x, y, z, l, m, n, ux, uy, uz)
    cb, cp = _in(coeff)
    r, a = _orders(rorder, aorder, cb.shape[0])
    lib().pxfo_tracezernopd(*p, num, cp, r.ctypes.data_as(_ip), a.ctypes.data_as(_ip), cb.shape[0], rad, nr)


def _zernphase(opd, x, y, z, l, m, n, ux, uy, uz, coeff, rorder, aorder, rad, wave):
    num, p = _io(opd, x, y, z, l, m, n, ux, uy, uz)
    cb, cp = _in(coeff)
    r, a = _orders(rorder, aorder, cb.shape[0])
    lib().pxfo_zernphase(*p, num, cp, r.ctypes.data_as(_ip), a.ctypes.data_as(_ip), cb.shape[0], rad, wave)


def _tracezernrot(x, y, z, l, m, n, ux, uy, uz, coeff1, rorder1, aorder1, coeff2, rorder2, aorder2, rad, rot):
    num, p = _io(x, y, z, l, m, n, ux, uy, uz)
    c1, c1p = _in(coeff1)
    r1, a1 = _orders(rorder1, aorder1, c1.shape[0])
    c2, c2p = _in(coeff2)
    r2, a2 = _orders(rorder2, aorder2, c2.shape[0])
    lib().pxfo_tracezernrot(*p, num, c1p, r1.ctypes.data_as(_ip), a1.ctypes.data_as(_ip), c1.shape[0],
                            c2p, r2.ctypes.data_as(_ip), a2.ctypes.data_as(_ip), c2.shape[0], rad, rot)


zernsurf = SimpleNamespace(tracezern=_tracezern, tracezernopd=_tracezernopd, zernphase=_zernphase,
                           tracezernrot=_tracezernrot)


# ---------------------------------------------------------------- specialfunctions
def _zernset(rho, theta, rorder, aorder):
    r = np.ascontiguousarray(rorder, dtype=np.int32)
    a = np.ascontiguousarray(aorder, dtype=np.int32)
    znum = r.shape[0]
    po, dr, dt = (np.zeros(znum) for _ in range(3))
    lib().pxfo_zernset(float(rho), float(theta), r.ctypes.data_as(_ip), a.ctypes.data_as(_ip), znum,
                       po.ctypes.data_as(_dp), dr.ctypes.data_as(_dp), dt.ctypes.data_as(_dp))
    return po, dr, dt


specialfunctions = SimpleNamespace(
    zernset=_zernset,
    radialpoly=lambda rho, n, m: lib().pxfo_radialpoly(float(rho), int(n), int(m)),
    legendre=lambda x, n: lib().pxfo_legendre(float(x), int(n)),
    legendrep=lambda x, n: lib().pxfo_legendrep(float(x), int(n)),
)


# ---------------------------------------------------------------- reconstruct (reconstruct.f95)
def _io2(*arrs):
    """intent(inout) 2-D arrays: f2py wants Fortran-contiguous float64 of one shape."""
    shape = None
    out = []
    for a in arrs:
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.ndim != 2 or not a.flags.f_contiguous:
            raise ValueError("failed in converting argument to C/Fortran array: "
                             "intent(inout) array must be a Fortran-contiguous 2-D float64 ndarray")
        if shape is None:
            shape = a.shape
        elif a.shape != shape:
            raise ValueError("shape mismatch against xdim,ydim")
        out.append(a.ctypes.data_as(_dp))
    return shape, out


def _reconstruct(xang, yang, criteria, h, phase, maxiter):
    """phasec = reconstruct(xang,yang,criteria,h,phase,maxiter) -- the f2py signature of reconstruct.f95:1
    (xdim, ydim hidden; xang, yang, phase intent(inout), phasec intent(out)).  ``_reconstruct.sweeps`` holds the
    number of sweeps of the last call (test aid)."""
    (xdim, ydim), (px, py, pp) = _io2(xang, yang, phase)
    phasec = np.zeros((xdim, ydim), order="F")
    _reconstruct.sweeps = int(lib().pxfo_reconstruct(px, py, xdim, ydim, float(criteria), float(h), pp,
                                                      phasec.ctypes.data_as(_dp), int(maxiter)))
    return phasec


def _southwellbin(x, y, l, m, binsize, xdim, ydim):
    """xang,yang,phase = southwellbin(x,y,l,m,binsize,xdim,ydim) (reconstruct.f95:136)."""
    xs, px = _in(x)
    n = xs.shape[0]
    ys, py = _in(y, n)
    ls, pl = _in(l, n)
    ms, pm = _in(m, n)
    xang, yang, phase = (np.zeros((int(xdim), int(ydim)), order="F") for _ in range(3))
    lib().pxfo_southwellbin(px, py, pl, pm, n, float(binsize), xang.ctypes.data_as(_dp), yang.ctypes.data_as(_dp),
                            phase.ctypes.data_as(_dp), int(xdim), int(ydim))
    return xang, yang, phase


reconstruct = SimpleNamespace(reconstruct=_reconstruct, southwellbin=_southwellbin)
