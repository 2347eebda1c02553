"""Mirror of the reference's ``southwell.py`` (padArrays :7-19, southwell :21-46) on the device reconstructor."""
import numpy as np

from . import reconstruct


def padArrays(imglist):
    """Pad arrays with a border of 100s, which the reconstructor interprets as NaNs (southwell.py:7-19)."""
    outimg = []
    for i in imglist:
        sh = np.shape(i)
        temp = np.zeros((sh[0] + 2, sh[1] + 2), order='F') + 100.
        temp[1:-1, 1:-1] = i
        outimg.append(temp)
    return outimg


def southwell(gx, gy, criteria, h, maxiter=10000):
    """Southwell reconstruction of a phase map from its x and y gradient arrays (southwell.py:21-46).  As in the
    reference the ``criteria`` and ``h`` arguments are NOT what reaches the reconstructor: it is called with
    1e-10 and 1. (:41).  ``gx``, ``gy`` are modified in place (missing data -> 100.)."""
    ind = np.logical_or(np.isnan(gx), np.isnan(gy))
    gx[ind] = 100.
    gy[ind] = 100.
    phase = np.zeros(np.shape(gx), order='F')
    phase[ind] = 100.
    phase, gx, gy = padArrays([phase, gx, gy])
    # the reference discards the returned phasec and keeps `phase`, which the Fortran updates in place
    # (intent(inout), reconstruct.f95:114)
    reconstruct.reconstruct(gx, gy, 1e-10, 1., phase, maxiter)
    phase = phase[1:-1, 1:-1]
    phase[ind] = np.nan
    return -phase
