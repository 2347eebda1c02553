"""Host-array entry: trace a bundle that lives in HOST memory (ten numpy arrays, the
reference's own representation) through a fused program on the device, in place.

This is the reference-facing drop-in for callers that keep numpy arrays: one call replaces a
whole sequence of f2py routine calls; the bundle streams through the GPU in chunks with the
PCIe transfers overlapped with the kernel (``pxf_host_trace_program``, include/pxf.h).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .program import Program

# z, l, m, n of subannulus / annulus / circularbeam (sources.py:56-170): one value for the whole bundle
SOURCE_CONST_ROWS = (3, 4, 5, 6)


def trace(rays, prog, write_back=True, hpd=False, alive=False, keep_xy=None, const_rows=()):
    """Run ``prog`` (a ``Program``) on a host bundle.

    rays : list of ten 1-D contiguous float64 numpy arrays (or torch CPU tensors, e.g.
           pinned), mutated in place like the Fortran does.  Entries the program neither
           reads nor writes may be None.
    keep_xy : optional pair of CUDA float64 tensors (length num) that receive the final x,y
           and stay resident on the device (e.g. for ``dist.hpd`` over a sharded bundle).
    const_rows : indices of input rows the caller knows to be constant over the bundle (z, l, m, n of every
           ``sources.*`` bundle: ``host.SOURCE_CONST_ROWS``); they are filled on the device instead of being
           scanned and uploaded.
    Returns a dict with ``hpd`` (if requested), ``alive`` (uint8 flags, if requested and the
    program vignettes) and ``alive_count``.
    """
    if not isinstance(prog, Program) or not len(prog):
        raise ValueError("need a non-empty Program")
    ptrs = []
    num = None
    keep = []
    for r in rays:
        if r is None:
            ptrs.append(None)
            continue
        if hasattr(r, "data_ptr"):          # torch CPU tensor (possibly pinned)
            if r.is_cuda or r.dtype != torch.float64 or r.dim() != 1 or not r.is_contiguous():
                raise ValueError("host rows must be contiguous 1-D float64 CPU arrays")
            n, p = r.shape[0], r.data_ptr()
        else:
            if not isinstance(r, np.ndarray) or r.dtype != np.float64 or r.ndim != 1 or not r.flags.c_contiguous:
                raise ValueError("host rows must be contiguous 1-D float64 CPU arrays")
            n, p = r.shape[0], r.ctypes.data
        if num is None:
            num = n
        elif n != num:
            raise ValueError("shape mismatch between rows")
        keep.append(r)
        ptrs.append(p)
    if num is None:
        raise ValueError("no rows given")
    tab = (ctypes.c_void_p * 10)(*ptrs)
    ops = prog.c_ops(torch.device("cuda", torch.cuda.current_device()) if prog._tables else None)
    h = ctypes.c_double(float("nan"))
    cnt = ctypes.c_int64(-1)
    flags = np.empty(num, dtype=np.uint8) if (alive and prog.has_vignette()) else None
    mask = 0
    for r in const_rows:
        mask |= 1 << int(r)
    rc = _lib.lib().pxf_host_trace_program_hint(tab, num, ops, len(prog), 1 if write_back else 0,
                                                ctypes.byref(h) if hpd else None,
                                                flags.ctypes.data if flags is not None else None,
                                                ctypes.byref(cnt),
                                                keep_xy[0].data_ptr() if keep_xy is not None else None,
                                                keep_xy[1].data_ptr() if keep_xy is not None else None, mask)
    _lib.check(rc)
    out = {"alive_count": int(cnt.value)}
    if hpd:
        out["hpd"] = float(h.value)
    if flags is not None:
        out["alive"] = flags
    return out
