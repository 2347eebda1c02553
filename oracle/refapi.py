"""The reference's Python API surface over the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

``load()`` returns a namespace shaped like ``oracle.refload.load()`` -- ``sources``, ``transformations``,
``surfaces``, ``analyses``, ``conicsolve`` with the reference's function names and argument order -- so that
one script text (``pyxfocus_b200/examples.py``: the BASELINE configurations written the way the reference's
example scripts are written) runs on three back ends:

    oracle.refload.load()        the reference's own unmodified Python layer + C oracle   (build container; goldens)
    oracle.refapi.load()         this file: numpy restatement of that layer + C oracle    (travels to the GPU box)
    pyxfocus_b200                the product                                              (what is being checked)

Only what the configurations call is restated; each function cites the reference lines it follows.
``tests/test_examples.py`` runs every configuration on refload and refapi and demands identical bits whenever
the reference tree is present.
"""
from types import SimpleNamespace

import numpy as np

from . import f2py as _f
from . import pyref

T, S, W, Z = _f.transformationsf, _f.surfacesf, _f.woltsurf, _f.zernsurf


def _gather(rows, ind):
    return [np.ascontiguousarray(r[ind]) for r in rows]


def _scatter(rows, ind, tmp):
    for r, t in zip(rows, tmp):
        r[ind] = t


# ------------------------------------------------------------------ transformations.py
def transform(rays, dx, dy, dz, rx, ry, rz, ind=None, coords=None):
    """transformations.py:11-44 (arguments negated :24,29; ind= gather/scatter :20-27)"""
    if coords is not None:
        raise NotImplementedError("refapi: coords bookkeeping is host-side numpy, not on the checked path")
    rows = rays[1:]
    if ind is not None:
        tmp = _gather(rows, ind)
        T.transform(*tmp, -dx, -dy, -dz, -rx, -ry, -rz)
        _scatter(rows, ind, tmp)
    else:
        T.transform(*rows, -dx, -dy, -dz, -rx, -ry, -rz)


def itransform(rays, dx, dy, dz, rx, ry, rz, coords=None, ind=None):
    """transformations.py:47-77"""
    if coords is not None:
        raise NotImplementedError("refapi: coords bookkeeping is host-side numpy, not on the checked path")
    rows = rays[1:]
    if ind is not None:
        tmp = _gather(rows, ind)
        T.itransform(*tmp, -dx, -dy, -dz, -rx, -ry, -rz)
        _scatter(rows, ind, tmp)
    else:
        T.itransform(*rows, -dx, -dy, -dz, -rx, -ry, -rz)


def reflect(rays, ind=None):
    """transformations.py:102-112"""
    rows = rays[4:]
    if ind is not None:
        tmp = _gather(rows, ind)
        T.reflect(*tmp)
        _scatter(rows, ind, tmp)
    else:
        T.reflect(*rows)


def radgrat(rays, dpermm, order, wave, ind=None):
    """transformations.py:124-172: ndarray wavelength -> radgratw, else radgrat; wave[ind] for masked arrays"""
    x, y, z, l, m, n = rays[1:7]
    fn = T.radgratw if type(wave) == np.ndarray else T.radgrat
    if ind is not None:
        tmp = _gather([x, y, l, m, n], ind)
        tw = wave if np.size(wave) == 1 else np.ascontiguousarray(wave[ind])
        fn(*tmp, tw, dpermm, order)
        _scatter([x, y, l, m, n], ind, tmp)
    else:
        fn(x, y, l, m, n, wave, dpermm, order)


vignette = pyref.vignette        # transformations.py:214-225


# ------------------------------------------------------------------ surfaces.py
def flat(rays, ind=None, nr=None):
    """surfaces.py:14-29 (ind= wins over nr=)"""
    if ind is not None:
        tmp = _gather(rays[1:], ind)
        S.flat(*tmp)
        _scatter(rays[1:], ind, tmp)
    elif nr is not None:
        S.flatopd(*rays[1:], rays[0], nr)
    else:
        S.flat(*rays[1:])


def zernsurf(rays, coeff, rad, rorder=None, aorder=None, nr=None):
    """surfaces.py:31-47 (the default ordering lives in an un-vendored module: orders are mandatory here)"""
    if rorder is None or aorder is None:
        raise NotImplementedError("zernsurf: pass rorder/aorder (utilities.imaging.zernikemod is not vendored)")
    c = np.ascontiguousarray(coeff, dtype=np.float64)
    if nr is None:
        Z.tracezern(*rays[1:], c, np.array(rorder), np.array(aorder), rad)
    else:
        Z.tracezernopd(*rays, c, np.array(rorder), np.array(aorder), rad, nr)


def wolterprimary(rays, r0, z0, psi=1., nr=None):
    """surfaces.py:219-227"""
    if nr is None:
        W.wolterprimary(*rays[1:], r0, z0, psi)
    else:
        W.wolterprimaryopd(*rays, r0, z0, psi, nr)


def woltersecondary(rays, r0, z0, psi=1.):
    """surfaces.py:238-243"""
    W.woltersecondary(*rays[1:], r0, z0, psi)


def wsPrimary(rays, r0, z0, psi, check=False):
    """surfaces.py:331-347 (check=True is broken in the reference; not restated)"""
    W.wsprimary(*rays[1:], pyref.woltparam(r0, z0)[0], z0, psi)


def wsSecondary(rays, r0, z0, psi, check=False):
    """surfaces.py:367-383"""
    W.wssecondary(*rays[1:], pyref.woltparam(r0, z0)[0], z0, psi)


def spoCone(rays, R0, tg, ind=None):
    """surfaces.py:403-419"""
    if ind is not None:
        tmp = _gather(rays[1:], ind)
        W.spocone(*tmp, R0, tg)
        _scatter(rays[1:], ind, tmp)
    else:
        W.spocone(*rays[1:], R0, tg)


def spoPrimary(rays, R0, F, d=.605, ind=None):
    """surfaces.py:421-430"""
    spoCone(rays, R0, .25 * np.arctan((R0 + d / 2) / F), ind=ind)


def spoSecondary(rays, R0, F, d=.605, ind=None):
    """surfaces.py:432-441"""
    spoCone(rays, R0, .75 * np.arctan((R0 + d / 2) / F), ind=ind)


def focus(rays, fn, weights=None, nr=None, coords=None):
    """surfaces.py:502-510"""
    dz1 = fn(rays, weights=weights)
    transform(rays, 0, 0, dz1, 0, 0, 0, coords=coords)
    flat(rays, nr=nr)
    dz2 = fn(rays, weights=weights)
    transform(rays, 0, 0, dz2, 0, 0, 0, coords=coords)
    flat(rays, nr=nr)
    return dz1 + dz2


def focusI(rays, weights=None, nr=None, coords=None):
    return focus(rays, pyref.analyticImagePlane, weights=weights, nr=nr, coords=coords)


def focusY(rays, weights=None, nr=None, coords=None):
    return focus(rays, analyticYPlane, weights=weights, nr=nr, coords=coords)


# ------------------------------------------------------------------ analyses.py
def analyticYPlane(rays, weights=None):
    """analyses.py:135-144"""
    x, y, z, l, m, n = rays[1:7]
    by = np.average(y * m / n, weights=weights) - np.average(y, weights=weights) * np.average(m / n, weights=weights)
    ay = np.average((m / n) ** 2, weights=weights) - np.average(m / n, weights=weights) ** 2
    return -by / ay


def rmsY(rays, weights=None):
    """analyses.py:53-58"""
    y = rays[2]
    cy = np.average(y, weights=weights)
    return np.sqrt(np.average((y - cy) ** 2, weights=weights))


def hpdY(rays, weights=None):
    """analyses.py:99-116"""
    y = rays[2]
    cy = np.average(y, weights=weights)
    rho = np.abs(y - cy)
    if weights is not None:
        ind = np.argsort(rho)
        weights = weights[ind]
        rho = rho[ind]
        cdf = np.cumsum(weights)
        cdf = cdf / cdf.max()
        return rho[np.argmin(np.abs(cdf - .75))] - rho[np.argmin(np.abs(cdf - .25))]
    return np.median(rho) * 2.


def indAngle(rays, ind=None, normal=None):
    """analyses.py:160-179 (current-normal form)"""
    if normal is not None:
        raise NotImplementedError
    r = rays if ind is None else [None] + [q[ind] for q in rays[1:]]
    return np.arccos(r[4] * r[7] + r[5] * r[8] + r[6] * r[9])


def grazeAngle(rays, ind=None):
    """analyses.py:181-184"""
    return np.pi / 2 - indAngle(rays, ind=ind)


def interpolateVec(rays, I, Nx, Ny, xr=None, yr=None, method='linear', polar=False, interpVec=None):
    """analyses.py:189-230 with scipy's own griddata (the reference's third-party dependency; scipy is present on the
    GPU box too, so this oracle travels)."""
    from scipy.interpolate import griddata
    x, y = rays[1:3]
    if interpVec is None:
        interpVec = rays[I]
    if xr is None:
        xr = [x.min(), x.max()]
        yr = [y.min(), y.max()]
    gridx, gridy = np.meshgrid(np.linspace(xr[0], xr[1], Nx), np.linspace(yr[0], yr[1], Ny))
    dx = np.diff(gridx)[0][0]
    dy = np.diff(np.transpose(gridy))[0][0]
    if polar is True:
        rho = np.sqrt(x ** 2 + y ** 2)
        theta1 = np.arctan2(y, x)
        theta2 = np.arctan2(x, y)
        rhog = np.sqrt(gridx ** 2 + gridy ** 2)
        azg1 = rhog * np.arctan2(gridy, gridx)
        azg2 = rhog * np.arctan2(gridx, gridy)
        res1 = griddata((rho, theta1 * rho), interpVec, (rhog, azg1), method=method)
        res2 = griddata((rho, theta2 * rho), interpVec, (rhog, azg2), method=method)
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = np.nanmedian([res1, res2], axis=0)
    else:
        res = griddata((x, y), interpVec, (gridx, gridy), method=method)
    return res, dx, dy


def wavefront(rays, Nx, Ny, method='cubic', polar=False, maxiter=10000):
    """analyses.py:305-334.  PARITY UNPINNED at two points the reference itself cannot execute: ``man.padRect`` is in the
    un-vendored ``utilities.imaging`` package (taken as a one-pixel NaN frame: the function strips ``[1:-1,1:-1]`` at
    the end) and ``reconstruct.reconstruct`` is called without its required ``maxiter`` (:327)."""
    y, dx, dy = interpolateVec(rays, 5, Nx, Ny, method=method, polar=polar)
    x, dx, dy = interpolateVec(rays, 4, Nx, Ny, method=method, polar=polar)

    def padRect(img):
        out = np.full((img.shape[0] + 2, img.shape[1] + 2), np.nan)
        out[1:-1, 1:-1] = img
        return out
    x = padRect(x)
    y = padRect(y)
    phase = np.zeros(np.shape(x), order='F')
    phase[np.isnan(x)] = 100.
    x[np.isnan(x)] = 100.
    y[np.isnan(y)] = 100.
    y = np.array(y, order='F')
    x = np.array(x, order='F')
    phase = _f.reconstruct.reconstruct(y, x, 1e-12, dx, phase, maxiter)
    phase[phase == 100] = np.nan
    x[x == 100] = np.nan
    y[y == 100] = np.nan
    return phase[1:-1, 1:-1], x[1:-1, 1:-1], y[1:-1, 1:-1]


# ------------------------------------------------------------------ conicsolve.py
def primrad(z, r0, z0, psi=1.):
    """conicsolve.py:7-15"""
    alpha = .25 * np.arctan(r0 / z0)
    thetah = (2 * (1 + 2 * psi)) / (1 + psi) * alpha
    thetap = (2 * psi) / (1 + psi) * alpha
    p = z0 * np.tan(4 * alpha) * np.tan(thetap)
    d = z0 * np.tan(4 * alpha) * np.tan(4 * alpha - thetah)
    e = np.cos(4 * alpha) * (1 + np.tan(4 * alpha) * np.tan(thetah))
    return np.sqrt(p ** 2 + 2 * p * z + (4 * e ** 2 * p * d) / (e ** 2 - 1))


def secrad(z, r0, z0, psi=1.):
    """conicsolve.py:29-37"""
    alpha = .25 * np.arctan(r0 / z0)
    thetah = (2 * (1 + 2 * psi)) / (1 + psi) * alpha
    thetap = (2 * psi) / (1 + psi) * alpha
    p = z0 * np.tan(4 * alpha) * np.tan(thetap)
    d = z0 * np.tan(4 * alpha) * np.tan(4 * alpha - thetah)
    e = np.cos(4 * alpha) * (1 + np.tan(4 * alpha) * np.tan(thetah))
    return np.sqrt(e ** 2 * (d + z) ** 2 - z ** 2)


def load():
    return SimpleNamespace(
        sources=SimpleNamespace(subannulus=pyref.subannulus, annulus=pyref.annulus, pointsource=pyref.pointsource,
                                circularbeam=pyref.circularbeam),
        transformations=SimpleNamespace(transform=transform, itransform=itransform, reflect=reflect, radgrat=radgrat,
                                        vignette=vignette),
        surfaces=SimpleNamespace(flat=flat, zernsurf=zernsurf, wolterprimary=wolterprimary,
                                 woltersecondary=woltersecondary, wsPrimary=wsPrimary, wsSecondary=wsSecondary,
                                 spoCone=spoCone, spoPrimary=spoPrimary, spoSecondary=spoSecondary, focus=focus,
                                 focusI=focusI, focusY=focusY),
        analyses=SimpleNamespace(centroid=pyref.centroid, rmsCentroid=pyref.rmsCentroid, rho=pyref.rho,
                                 rhocdf=pyref.rhocdf, hpd=pyref.hpd, analyticImagePlane=pyref.analyticImagePlane,
                                 analyticYPlane=analyticYPlane, rmsY=rmsY, hpdY=hpdY, indAngle=indAngle,
                                 grazeAngle=grazeAngle),
        conicsolve=SimpleNamespace(primrad=primrad, secrad=secrad, woltparam=pyref.woltparam))
