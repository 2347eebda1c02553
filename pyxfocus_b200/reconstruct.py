"""Replacement for the reference's f2py module ``reconstruct`` (reconstruct.f95; imported by
``analyses.py:11`` and ``southwell.py:3``): same routine names and f2py signatures.

    phasec = reconstruct(xang, yang, criteria, h, phase, maxiter)        # reconstruct.f95:1-128
    xang, yang, phase = southwellbin(x, y, l, m, binsize, xdim, ydim)    # reconstruct.f95:136-187

``xang``, ``yang``, ``phase`` are the reference's Fortran-ordered 2-D float64 numpy arrays (``intent(inout)``:
mutated in place, anything else raises ``ValueError`` as the f2py wrapper does); they are staged through the
device, where the Gauss-Seidel sweeps run as a diagonal-parity pipeline (``csrc/pxf_reconstruct.cu``) --
bit-identical to the sequential Fortran loop.  ``southwellbin`` takes the bundle rows where they live (CUDA
tensors or numpy arrays)."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._call import stream_ptr


class error(Exception):
    pass


def _device():
    if not torch.cuda.is_available():
        raise _lib.PxfError("pyxfocus_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _inout2(*arrs):
    shape = None
    for a in arrs:
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.ndim != 2 or not a.flags.f_contiguous:
            raise ValueError("failed in converting argument to C/Fortran array: intent(inout) array must be a "
                             "Fortran-contiguous 2-D float64 ndarray")
        if shape is None:
            shape = a.shape
        elif a.shape != shape:
            raise ValueError("shape mismatch against xdim,ydim")
    return shape


def _up(a, dev):
    # column-major [xdim][ydim] == the row-major memory of the transposed array
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)


def reconstruct(xang, yang, criteria, h, phase, maxiter, xdim=None, ydim=None):
    """Southwell reconstruction by successive over-relaxation; returns ``phasec``.  ``reconstruct.sweeps`` holds
    the number of sweeps of the last call.  ``xdim``, ``ydim``: f2py's optional trailing shape arguments."""
    shape = _inout2(xang, yang, phase)
    if (xdim is not None and int(xdim) != shape[0]) or (ydim is not None and int(ydim) != shape[1]):
        raise ValueError("shape(xang,0)==xdim / shape(xang,1)==ydim failed")
    xdim, ydim = shape
    dev = _device()
    dx, dy, dp = _up(xang, dev), _up(yang, dev), _up(phase, dev)
    dc = torch.empty_like(dp)
    sweeps = ctypes.c_int64(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pxf_reconstruct(dx.data_ptr(), dy.data_ptr(), xdim, ydim, float(criteria), float(h),
                                              dp.data_ptr(), dc.data_ptr(), int(maxiter), ctypes.byref(sweeps),
                                              stream_ptr(dev)))
    xang[...] = dx.cpu().numpy().T
    yang[...] = dy.cpu().numpy().T
    phase[...] = dp.cpu().numpy().T
    reconstruct.sweeps = int(sweeps.value)
    return np.asfortranarray(dc.cpu().numpy().T)


def southwellbin(x, y, l, m, binsize, xdim, ydim):
    """Bin ray positions / direction cosines into a lenslet array; returns Fortran-ordered ``xang, yang, phase``."""
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else _device()
    rows = [torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v, dtype=torch.float64).to(dev).contiguous()
            for v in (x, y, l, m)]
    num = rows[0].shape[0]
    if any(r.shape[0] != num for r in rows):
        raise ValueError("shape mismatch against num")
    xdim, ydim = int(xdim), int(ydim)
    L = _lib.lib()
    out = torch.empty(3, ydim, xdim, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        scratch = torch.empty(int(L.pxf_southwellbin_scratch_bytes(num, xdim, ydim)), dtype=torch.uint8, device=dev)
        _lib.check(L.pxf_southwellbin(rows[0].data_ptr(), rows[1].data_ptr(), rows[2].data_ptr(), rows[3].data_ptr(), num,
                                      float(binsize), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), xdim, ydim,
                                      scratch.data_ptr(), stream_ptr(dev)))
    h = out.cpu().numpy()
    return tuple(np.asfortranarray(h[k].T) for k in range(3))
