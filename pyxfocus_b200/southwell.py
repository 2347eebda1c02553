"""Southwell reconstruction driver on the device reconstructor: what the reference's ``southwell.py`` provides
(``padArrays`` :7-19, ``southwell`` :21-46)."""
import numpy as np

from . import reconstruct

MISSING = 100.      # the reconstructor's marker for "no data" (reconstruct.f95 treats 100. as NaN)


def padArrays(imglist):
    """Every image framed by one pixel of ``MISSING`` (southwell.py:7-19); Fortran-ordered copies."""
    return [np.asfortranarray(np.pad(np.asarray(img, dtype=np.float64), 1, constant_values=MISSING)) for img in imglist]


def southwell(gx, gy, criteria, h, maxiter=10000):
    """Phase map from its x and y slope maps (southwell.py:21-46).  Like the reference, the ``criteria`` and ``h``
    arguments do NOT reach the reconstructor (it is called with 1e-10 and 1., :41), missing slopes are overwritten
    with ``MISSING`` in the caller's arrays, and the result is the NEGATED in-place ``phase`` (the returned ``phasec``
    of the Fortran is dropped, reconstruct.f95:114)."""
    missing = np.isnan(gx) | np.isnan(gy)
    gx[missing] = MISSING
    gy[missing] = MISSING
    seed = np.where(missing, MISSING, 0.)
    phase, px, py = padArrays([seed, gx, gy])
    reconstruct.reconstruct(px, py, 1e-10, 1., phase, maxiter)
    out = phase[1:-1, 1:-1]
    out[missing] = np.nan
    return -out
