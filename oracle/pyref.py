"""numpy restatement of the reference's *Python* arithmetic on the hot path
(TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

The reference's Python layer cannot travel to the GPU box (``/root/reference`` does not exist
there), so the handful of numpy formulas that sit on the path -- sources, vignette, the
analyses -- are restated here, each citing the lines it follows.  ``tests/test_oracle_golden``
checks every function against the real reference Python (imported unmodified through
``oracle.refload``) whenever the reference tree is present, and against the committed golden
vectors otherwise.
"""
import numpy as np

from . import f2py as _f


# ---------------------------------------------------------------- sources.py
def _zeros(num):
    return np.repeat(0., num)


def subannulus(rin, rout, dphi, num, zhat=1.):
    """sources.py:130-170 -- radius vector drawn first, then the angle vector (:157-158)."""
    rho = np.sqrt(rin ** 2 + np.random.rand(num) * (rout ** 2 - rin ** 2))
    theta = np.random.rand(num) * dphi - dphi / 2.
    x = rho * np.cos(theta)
    y = rho * np.sin(theta)
    z, l, m = _zeros(num), _zeros(num), _zeros(num)
    n = np.repeat(zhat, num)
    return [_zeros(num), x, y, z, l, m, n, _zeros(num), _zeros(num), _zeros(num)]


def annulus(rin, rout, num, zhat=-1.):
    """sources.py:91-127"""
    rho = np.sqrt(rin ** 2 + np.random.rand(num) * (rout ** 2 - rin ** 2))
    theta = np.random.rand(num) * 2 * np.pi
    x = rho * np.cos(theta)
    y = rho * np.sin(theta)
    return [_zeros(num), x, y, _zeros(num), _zeros(num), _zeros(num), np.repeat(zhat, num),
            _zeros(num), _zeros(num), _zeros(num)]


def circularbeam(rad, num):
    """sources.py:56-88"""
    rho = np.sqrt(np.random.rand(num)) * rad
    theta = np.random.rand(num) * 2 * np.pi
    x = rho * np.cos(theta)
    y = rho * np.sin(theta)
    return [_zeros(num), x, y, _zeros(num), _zeros(num), _zeros(num), np.repeat(1., num),
            _zeros(num), _zeros(num), _zeros(num)]


def pointsource(ang, num):
    """sources.py:20-53"""
    rho = np.sqrt(np.random.rand(num)) * np.sin(ang)
    theta = np.random.rand(num) * 2 * np.pi
    l = rho * np.cos(theta)
    m = rho * np.sin(theta)
    n = np.sqrt(1. - l ** 2 - m ** 2)
    return [_zeros(num), _zeros(num), _zeros(num), _zeros(num), l, m, n, _zeros(num), _zeros(num), _zeros(num)]


# ---------------------------------------------------------------- sources.py, the set-up sources (:173-471)
# numpy restatements (same expressions, same order of draws from numpy's global stream); the product generates
# these on the device (pxf_source_grid / pxf_source_beam) and the GPU tests compare against these.
def _bundle(opd, x, y, z, l, m, n, ux, uy, uz):
    return [np.array(a, dtype=np.float64) for a in (opd, x, y, z, l, m, n, ux, uy, uz)]


def xslit(xin, xout, num, zhat=-1.):
    """Slit of rays linearly spaced in x (sources.py:173-207)."""
    x = np.linspace(xin, xout, num)
    zero = np.repeat(0., num)
    return _bundle(zero, x, zero, zero, zero, zero, np.repeat(zhat, num), zero, zero, zero)


def rectArray(xsize, ysize, num):
    """num x num rectangular grid of rays in +z (sources.py:210-247)."""
    x, y = np.meshgrid(np.linspace(-xsize, xsize, num), np.linspace(-ysize, ysize, num))
    zero = np.repeat(0., num ** 2)
    return _bundle(zero, x.flatten(), y.flatten(), zero, zero, zero, np.repeat(1., num ** 2), zero, zero, zero)


def _converging(x, y, rho, theta, zset, num, lscat):
    z = np.repeat(zset, num)
    lscat = lscat * np.tan((np.random.rand(num) - .5) * np.pi)
    lscat = lscat / 60 ** 2 * np.pi / 180.
    n = -np.cos(np.arctan(rho / zset) + lscat)
    l = -np.sqrt(1 - n ** 2) * np.cos(theta)
    m = -np.sqrt(1 - n ** 2) * np.sin(theta)
    zero = np.repeat(0., num)
    return _bundle(zero, x, y, z, l, m, n, zero, zero, zero)


def convergingbeam(zset, rin, rout, tmin, tmax, num, lscat):
    """Converging sub-apertured annulus beam placed at its nominal focus (sources.py:250-296)."""
    rho = np.sqrt(rin ** 2 + np.random.rand(num) * (rout ** 2 - rin ** 2))
    theta = tmin + np.random.rand(num) * (tmax - tmin)
    x = rho * np.cos(theta)
    y = rho * np.sin(theta)
    return _converging(x, y, rho, theta, zset, num, lscat)


def convergingbeam2(zset, xmin, xmax, ymin, ymax, num, lscat):
    """Converging rectangular beam placed at its nominal focus (sources.py:299-345)."""
    x = xmin + np.random.rand(num) * (xmax - xmin)
    y = ymin + np.random.rand(num) * (ymax - ymin)
    rho = np.sqrt(x ** 2 + y ** 2)
    theta = np.arctan2(y, x)
    return _converging(x, y, rho, theta, zset, num, lscat)


def rectbeam(xhalfwidth, yhalfwidth, num):
    """Uniform rectangular beam in +z (sources.py:348-379)."""
    x = (np.random.rand(num) - .5) * 2 * xhalfwidth
    y = (np.random.rand(num) - .5) * 2 * yhalfwidth
    zero = np.repeat(0., num)
    return _bundle(zero, x, y, zero, zero, zero, np.repeat(1., num), zero, zero, zero)


def gaussianBeam(ang, num):
    """Point source with a Gaussian angular profile (sources.py:381-416)."""
    l = np.random.randn(num) * np.sin(ang) / np.sqrt(2)
    m = np.random.randn(num) * np.sin(ang) / np.sqrt(2)
    n = np.sqrt(1. - l ** 2 - m ** 2)
    zero = np.repeat(0., num)
    return _bundle(zero, zero, zero, zero, l, m, n, zero, zero, zero)


def _fan(xa, ya):
    num = np.size(xa)
    l = np.sin(xa)
    m = np.sin(ya)
    n = np.sqrt(1. - l ** 2 - m ** 2)
    zero = np.repeat(0., num)
    return _bundle(zero, zero, zero, zero, l, m, n, zero, zero, zero)


def fanBeam(xang, yang, num):
    """Rectangular fan of rays from a point (sources.py:418-442)."""
    xa, ya = np.meshgrid(np.linspace(-xang, xang, num), np.linspace(-yang, yang, num))
    return _fan(xa.flatten(), ya.flatten())


def circFan(halfang, rings, arms):
    """Circular fan of rays from a point: ``rings`` radii x ``arms`` azimuths (sources.py:444-471)."""
    rad = np.linspace(0, halfang, rings)
    az = np.linspace(0, 2 * np.pi, arms + 1)[0:-1]
    rr, aa = np.meshgrid(rad, az)
    xx = np.sin(rr) * np.cos(aa)
    yy = np.sin(rr) * np.sin(aa)
    return _fan(xx.flatten(), yy.flatten())


# ---------------------------------------------------------------- transformations.py
def vignette(rays, ind=None):
    """transformations.py:214-225"""
    if ind is None:
        mag = rays[4] ** 2 + rays[5] ** 2 + rays[6] ** 2
        ind = np.where(mag > .1)
    return [rays[i][ind] for i in range(10)]


def transform(rays, dx, dy, dz, rx, ry, rz):
    """transformations.py:29 (arguments negated before the Fortran call)"""
    _f.transformationsf.transform(*rays[1:], -dx, -dy, -dz, -rx, -ry, -rz)


def itransform(rays, dx, dy, dz, rx, ry, rz):
    """transformations.py:63"""
    _f.transformationsf.itransform(*rays[1:], -dx, -dy, -dz, -rx, -ry, -rz)


def masked(fn, arrays, ind, *scalars):
    """The reference's ind= idiom: gather -> Fortran -> scatter (transformations.py:20-27)."""
    tmp = [np.ascontiguousarray(a[ind]) for a in arrays]
    fn(*tmp, *scalars)
    for a, t in zip(arrays, tmp):
        a[ind] = t


# ---------------------------------------------------------------- analyses.py
def centroid(rays, weights=None):
    """analyses.py:16-22"""
    return np.average(rays[1], weights=weights), np.average(rays[2], weights=weights)


def rmsCentroid(rays, weights=None):
    """analyses.py:24-30"""
    cx, cy = centroid(rays, weights=weights)
    rho = (rays[1] - cx) ** 2 + (rays[2] - cy) ** 2
    return np.sqrt(np.average(rho, weights=weights))


def rho(rays, weights=None, cent=False):
    """analyses.py:60-71"""
    if cent is True:
        cx, cy = centroid(rays, weights=weights)
    else:
        cx, cy = 0, 0
    return np.sqrt((rays[1] - cx) ** 2 + (rays[2] - cy) ** 2)


def rhocdf(rays, weights=None, cent=True):
    """analyses.py:73-86"""
    r = rho(rays, weights=weights, cent=cent)
    if weights is None:
        weights = np.repeat(1, len(r))
    ind = np.argsort(r)
    weights = weights[ind]
    r = r[ind]
    cdf = np.cumsum(weights)
    cdf = cdf / cdf.max()
    return r, cdf


def hpd(rays, weights=None):
    """analyses.py:88-97 (the cent argument is ignored there: always centroid-relative)"""
    r = rho(rays, weights=weights, cent=True)
    if weights is not None:
        r, cdf = rhocdf(rays, weights=weights, cent=True)
        return r[np.argmin(np.abs(cdf - .75))] - r[np.argmin(np.abs(cdf - .25))]
    return np.median(r) * 2.


def analyticImagePlane(rays, weights=None):
    """analyses.py:118-133"""
    x, y, z, l, m, n = rays[1:7]
    av = lambda q: np.average(q, weights=weights)   # noqa: E731
    bx = av(x * l / n) - av(x) * av(l / n)
    ax = av((l / n) ** 2) - av(l / n) ** 2
    by = av(y * m / n) - av(y) * av(m / n)
    ay = av((m / n) ** 2) - av(m / n) ** 2
    return -(bx + by) / (ax + ay)


def focusI(rays, weights=None):
    """surfaces.py:502-519 (two-pass best focus)"""
    dz1 = analyticImagePlane(rays, weights=weights)
    transform(rays, 0, 0, dz1, 0, 0, 0)
    _f.surfacesf.flat(*rays[1:])
    dz2 = analyticImagePlane(rays, weights=weights)
    transform(rays, 0, 0, dz2, 0, 0, 0)
    _f.surfacesf.flat(*rays[1:])
    return dz1 + dz2


def woltparam(r0, z0):
    """conicsolve.py:51-59"""
    alpha = .25 * np.arctan(r0 / z0)
    thetah = 3 * alpha
    thetap = alpha
    p = z0 * np.tan(4 * alpha) * np.tan(thetap)
    d = z0 * np.tan(4 * alpha) * np.tan(4 * alpha - thetah)
    e = np.cos(4 * alpha) * (1 + np.tan(4 * alpha) * np.tan(thetah))
    return alpha, p, d, e
