"""Ray sources, mirroring the reference's ``sources.py`` API on the device bundle.

A bundle is a Python list of ten 1-D float64 CUDA tensors ``[opd,x,y,z,l,m,n,ux,uy,uz]``
(rows of one ``[10,N]`` allocation), the direct analogue of the reference's list of ten
numpy arrays (sources.py:1-15).

"Identical ray seeds": with ``rng='numpy'`` (default) the two uniform vectors are drawn on the
host from numpy's legacy global MT19937 stream exactly as the reference does -- radius vector
first, then angle vector (sources.py:157-158) -- uploaded, and the geometry is evaluated by a
CUDA kernel.  ``rng='philox'`` draws on the device (counter-based Philox4x32-10 keyed by
``seed`` and the global ray index) for bundles too large for the host generator.

The remaining sources of the reference (``xslit``, ``rectArray``, ``convergingbeam``,
``convergingbeam2``, ``rectbeam``, ``gaussianBeam``, ``fanBeam``, ``circFan``; sources.py:173-471)
are device kernels too (``pxf_source_grid`` / ``pxf_source_beam``): the grid sources restate
``numpy.linspace`` / ``meshgrid`` index arithmetic, the beam sources take their draws either from
numpy's global stream (uploaded, in the reference's order of draws) or from Philox.
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


def _make(kind, num, a, b, c, d, rng, seed, first, device, uniforms=None, out=None):
    num = int(num)
    if out is not None:
        # generate into an existing bundle (e.g. one shell's segment of a nested assembly)
        rays = out
        dev = rays[1].device
        for r in rays:
            if not r.is_cuda or r.dtype != torch.float64 or not r.is_contiguous() or r.shape[0] != num:
                raise ValueError("out must be ten contiguous float64 CUDA rows of length num")
    else:
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


def pointsource(ang, num, rng="numpy", seed=0, first=0, device=None, uniforms=None, out=None):
    """Point source with half-angle ``ang``; rays point in +z (sources.py:20-53)."""
    return _make("pointsource", num, float(ang), 0., 0., 0., rng, seed, first, device, uniforms, out)


def circularbeam(rad, num, rng="numpy", seed=0, first=0, device=None, uniforms=None, out=None):
    """Uniform circular beam of radius ``rad``, rays in +z (sources.py:56-88)."""
    return _make("circularbeam", num, float(rad), 0., 0., 0., rng, seed, first, device, uniforms, out)


def annulus(rin, rout, num, zhat=-1., rng="numpy", seed=0, first=0, device=None, uniforms=None, out=None):
    """Annulus of rays (sources.py:91-127)."""
    return _make("annulus", num, float(rin), float(rout), 0., float(zhat), rng, seed, first, device, uniforms, out)


def subannulus(rin, rout, dphi, num, zhat=1., rng="numpy", seed=0, first=0, device=None, uniforms=None, out=None):
    """Sub-apertured annulus centred on theta=0 (+x) (sources.py:130-170)."""
    return _make("subannulus", num, float(rin), float(rout), float(dphi), float(zhat), rng, seed, first, device,
                 uniforms, out)


def segments(kind, params, sizes, seed=0, first=0, device=None, out=None):
    """A nested assembly's source in ONE launch: segment k holds ``sizes[k]`` rays of source ``kind``
    ('subannulus' (rin, rout, dphi, zhat), 'annulus' (rin, rout, 0, zhat) or 'circularbeam' (rad, 0, 0, 0)) with
    parameters ``params[k]``; device Philox stream, identical to per-segment calls with
    ``first + sum(sizes[:k])``.  Returns the bundle (``bundle_split(bundle, sizes)`` gives per-segment views)."""
    if kind not in ("subannulus", "annulus", "circularbeam"):
        raise ValueError("segments: kind must be subannulus, annulus or circularbeam")
    sizes = [int(v) for v in sizes]
    total = sum(sizes)
    par = np.zeros((len(sizes), 4), dtype=np.float64)
    for k, p in enumerate(params):
        par[k, :len(p)] = p
    start = np.zeros(len(sizes) + 1, dtype=np.int64)
    start[1:] = np.cumsum(sizes)
    if out is not None:
        rays, dev = out, out[1].device
        if rays[1].shape[0] != total:
            raise ValueError("out has %d rays, the segments add up to %d" % (rays[1].shape[0], total))
    else:
        dev = _device(device)
        rays = bundle_alloc(total, dev)
    ptrs = (ctypes.c_void_p * 10)(*[r.data_ptr() for r in rays])
    with torch.cuda.device(dev):
        st = torch.from_numpy(start).to(dev)
        pr = torch.from_numpy(par).to(dev)
        _lib.check(_lib.lib().pxf_source_segmented(_KIND[kind], ptrs, total, int(first), int(seed) & (2 ** 64 - 1),
                                                   len(sizes), st.data_ptr(), pr.data_ptr(), stream_ptr(dev)))
        st.record_stream(torch.cuda.current_stream(dev))
        pr.record_stream(torch.cuda.current_stream(dev))
    return rays


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


# ---- the reference's set-up sources (sources.py:173-471), generated on the device --------------------
_XSLIT, _RECTARRAY, _CONVERGING, _CONVERGING2, _RECTBEAM, _GAUSSIAN, _FANBEAM, _CIRCFAN = 4, 5, 6, 7, 8, 9, 10, 11


def _grid(kind, total, n1, n2, a, b, c, first, num, device, out):
    total = int(total)
    first = int(first)
    num = total - first if num is None else int(num)
    if out is not None:
        rays, dev = out, out[1].device
        if rays[1].shape[0] != num:
            raise ValueError("out has %d rays, the source (shard) has %d" % (rays[1].shape[0], num))
    else:
        dev = _device(device)
        rays = bundle_alloc(num, dev)
    ptrs = (ctypes.c_void_p * 10)(*[r.data_ptr() for r in rays])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pxf_source_grid(kind, ptrs, num, first, int(n1), int(n2), float(a), float(b), float(c),
                                              stream_ptr(dev)))
    return rays


def _beam(kind, num, par, ndraw, rng, seed, first, device, draws, out, normal=False):
    num = int(num)
    if out is not None:
        rays, dev = out, out[1].device
        if rays[1].shape[0] != num:
            raise ValueError("out has %d rays, asked for %d" % (rays[1].shape[0], num))
    else:
        dev = _device(device)
        rays = bundle_alloc(num, dev)
    ptrs = (ctypes.c_void_p * 10)(*[r.data_ptr() for r in rays])
    cpar = (ctypes.c_double * 6)(*([float(v) for v in par] + [0.] * (6 - len(par))))
    L = _lib.lib()
    with torch.cuda.device(dev):
        if rng == "philox":
            rc = L.pxf_source_beam(kind, ptrs, num, int(first), int(seed) & (2 ** 64 - 1), cpar, stream_ptr(dev))
        elif rng == "numpy":
            if draws is None:
                # numpy's global legacy stream, one vector per draw in the reference's order
                draws = [(np.random.randn if normal else np.random.rand)(num) for _ in range(ndraw)]
            if len(draws) != ndraw:
                raise ValueError("this source takes %d draw vectors" % ndraw)
            t = [torch.from_numpy(np.ascontiguousarray(d, dtype=np.float64)).to(dev) for d in draws]
            for d in t:
                if d.shape[0] != num:
                    raise ValueError("draw vectors must have num elements")
            rc = L.pxf_source_beam_from_draws(kind, ptrs, num, t[0].data_ptr(), t[1].data_ptr(),
                                              t[2].data_ptr() if ndraw > 2 else None, cpar, stream_ptr(dev))
            torch.cuda.current_stream(dev).synchronize()   # the uploaded draws are freed on return
        else:
            raise ValueError("rng must be 'numpy' or 'philox'")
    _lib.check(rc)
    return rays


def xslit(xin, xout, num, zhat=-1., first=0, count=None, device=None, out=None):
    """Slit of rays linearly spaced in x (sources.py:173-207).  ``first``/``count`` generate a shard."""
    return _grid(_XSLIT, num, num, 0, xin, xout, zhat, first, count, device, out)


def rectArray(xsize, ysize, num, first=0, count=None, device=None, out=None):
    """num x num rectangular grid of rays in +z (sources.py:210-247)."""
    return _grid(_RECTARRAY, int(num) ** 2, num, 0, xsize, ysize, 0., first, count, device, out)


def convergingbeam(zset, rin, rout, tmin, tmax, num, lscat, rng="numpy", seed=0, first=0, device=None, draws=None,
                   out=None):
    """Converging sub-apertured annulus beam placed at its nominal focus (sources.py:250-296).  With
    ``rng='numpy'`` the three uniform vectors (radius, angle, scatter) come from numpy's global stream in the
    reference's order -- or from ``draws`` -- and the geometry is evaluated on the device."""
    return _beam(_CONVERGING, num, (zset, rin, rout, tmin, tmax, lscat), 3, rng, seed, first, device, draws, out)


def convergingbeam2(zset, xmin, xmax, ymin, ymax, num, lscat, rng="numpy", seed=0, first=0, device=None,
                    draws=None, out=None):
    """Converging rectangular beam placed at its nominal focus (sources.py:299-345); draws: x, y, scatter."""
    return _beam(_CONVERGING2, num, (zset, xmin, xmax, ymin, ymax, lscat), 3, rng, seed, first, device, draws, out)


def rectbeam(xhalfwidth, yhalfwidth, num, rng="numpy", seed=0, first=0, device=None, draws=None, out=None):
    """Uniform rectangular beam in +z (sources.py:348-379); draws: x, y."""
    return _beam(_RECTBEAM, num, (xhalfwidth, yhalfwidth), 2, rng, seed, first, device, draws, out)


def gaussianBeam(ang, num, rng="numpy", seed=0, first=0, device=None, draws=None, out=None):
    """Point source with a Gaussian angular profile (sources.py:381-416); draws: two standard-normal vectors
    (``np.random.randn`` on the host, a Box-Muller pair per ray with ``rng='philox'``)."""
    return _beam(_GAUSSIAN, num, (ang,), 2, rng, seed, first, device, draws, out, normal=True)


def fanBeam(xang, yang, num, first=0, count=None, device=None, out=None):
    """Rectangular fan of rays from a point (sources.py:418-442)."""
    return _grid(_FANBEAM, int(num) ** 2, num, 0, xang, yang, 0., first, count, device, out)


def circFan(halfang, rings, arms, first=0, count=None, device=None, out=None):
    """Circular fan of rays from a point: ``rings`` radii x ``arms`` azimuths (sources.py:444-471)."""
    return _grid(_CIRCFAN, int(rings) * int(arms), rings, arms, halfang, 0., 0., first, count, device, out)
