"""Ray sources, mirroring the reference's ``sources.py`` API on the device bundle.

A bundle is a Python list of ten 1-D float64 CUDA tensors ``[opd,x,y,z,l,m,n,ux,uy,uz]``
(rows of one ``[10,N]`` allocation), the direct analogue of the reference's list of ten
numpy arrays (sources.py:1-15).

"Identical ray seeds": with ``rng='numpy'`` (default) the two uniform vectors are drawn on the
host from numpy's legacy global MT19937 stream exactly as the reference does -- radius vector
first, then angle vector (sources.py:157-158) -- uploaded, and the geometry is evaluated by a
CUDA kernel.  ``rng='philox'`` draws on the device (counter-based Philox4x32-10 keyed by
``seed`` and the global ray index) for bundles too large for the host generator.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._call import bundle_alloc, bundle_split, stream_ptr  # noqa: F401

_KIND = dict(subannulus=0, circularbeam=1, pointsource=2, annulus=3)


def _device(device):
    if not torch.cuda.is_available():
        raise _lib.PxfError("pyxfocus_b200 needs a CUDA device (there is no CPU fallback)")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _make(kind, num, a, b, c, d, rng, seed, first, device, uniforms=None):
    num = int(num)
    dev = _device(device)
    rays = bundle_alloc(num, dev)
    ptrs = (ctypes.c_void_p * 10)(*[r.data_ptr() for r in rays])
    L = _lib.lib()
    with torch.cuda.device(dev):
        if rng == "philox":
            rc = L.pxf_source(_KIND[kind], ptrs, num, int(first), int(seed) & (2 ** 64 - 1), a, b, c, d,
                              stream_ptr(dev))
        elif rng == "numpy":
            if uniforms is None:
                u1 = np.random.rand(num)      # radius first ...
                u2 = np.random.rand(num)      # ... then angle (sources.py:157-158)
            else:
                u1, u2 = uniforms
            t1 = torch.from_numpy(np.ascontiguousarray(u1, dtype=np.float64)).to(dev)
            t2 = torch.from_numpy(np.ascontiguousarray(u2, dtype=np.float64)).to(dev)
            rc = L.pxf_source_from_uniform(_KIND[kind], ptrs, num, t1.data_ptr(), t2.data_ptr(), a, b, c, d,
                                           stream_ptr(dev))
            torch.cuda.current_stream(dev).synchronize()   # t1/t2 are freed on return
        else:
            raise ValueError("rng must be 'numpy' or 'philox'")
    _lib.check(rc)
    return rays


def pointsource(ang, num, rng="numpy", seed=0, first=0, device=None, uniforms=None):
    """Point source with half-angle ``ang``; rays point in +z (sources.py:20-53)."""
    return _make("pointsource", num, float(ang), 0., 0., 0., rng, seed, first, device, uniforms)


def circularbeam(rad, num, rng="numpy", seed=0, first=0, device=None, uniforms=None):
    """Uniform circular beam of radius ``rad``, rays in +z (sources.py:56-88)."""
    return _make("circularbeam", num, float(rad), 0., 0., 0., rng, seed, first, device, uniforms)


def annulus(rin, rout, num, zhat=-1., rng="numpy", seed=0, first=0, device=None, uniforms=None):
    """Annulus of rays (sources.py:91-127)."""
    return _make("annulus", num, float(rin), float(rout), 0., float(zhat), rng, seed, first, device, uniforms)


def subannulus(rin, rout, dphi, num, zhat=1., rng="numpy", seed=0, first=0, device=None, uniforms=None):
    """Sub-apertured annulus centred on theta=0 (+x) (sources.py:130-170)."""
    return _make("subannulus", num, float(rin), float(rout), float(dphi), float(zhat), rng, seed, first, device,
                 uniforms)


def from_numpy(rays, device=None):
    """Upload a reference-style bundle (ten numpy arrays) into one device allocation."""
    dev = _device(device)
    num = int(np.shape(rays[1])[0])
    out = bundle_alloc(num, dev)
    for i in range(10):
        out[i].copy_(torch.from_numpy(np.ascontiguousarray(rays[i], dtype=np.float64)))
    return out


def to_numpy(rays):
    """Download a device bundle to the reference's representation."""
    from .program import flush
    flush(rays)
    return [r.detach().cpu().numpy() for r in rays]
