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
are set-up helpers, not part of the trace path: they evaluate the reference's numpy formulas on the
host -- same expressions, same order of draws from numpy's global stream -- and upload the bundle
once (``from_numpy``).
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


# ---- remaining reference sources: host numpy formulas, one upload -------------------------------
def _uploaded(fn):
    """``fn`` builds the reference-style bundle (ten numpy arrays) on the host; the public function
    uploads it.  ``<source>.host(...)`` is the host half alone (what the CPU tests compare with the
    reference)."""
    def wrapper(*args, device=None, **kwargs):
        return from_numpy(fn(*args, **kwargs), device=device)
    wrapper.host = fn
    wrapper.__name__ = fn.__name__
    wrapper.__doc__ = fn.__doc__
    return wrapper


def _bundle(opd, x, y, z, l, m, n, ux, uy, uz):
    return [np.array(a, dtype=np.float64) for a in (opd, x, y, z, l, m, n, ux, uy, uz)]


@_uploaded
def xslit(xin, xout, num, zhat=-1.):
    """Slit of rays linearly spaced in x (sources.py:173-207)."""
    x = np.linspace(xin, xout, num)
    zero = np.repeat(0., num)
    return _bundle(zero, x, zero, zero, zero, zero, np.repeat(zhat, num), zero, zero, zero)


@_uploaded
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


@_uploaded
def convergingbeam(zset, rin, rout, tmin, tmax, num, lscat):
    """Converging sub-apertured annulus beam placed at its nominal focus (sources.py:250-296)."""
    rho = np.sqrt(rin ** 2 + np.random.rand(num) * (rout ** 2 - rin ** 2))
    theta = tmin + np.random.rand(num) * (tmax - tmin)
    x = rho * np.cos(theta)
    y = rho * np.sin(theta)
    return _converging(x, y, rho, theta, zset, num, lscat)


@_uploaded
def convergingbeam2(zset, xmin, xmax, ymin, ymax, num, lscat):
    """Converging rectangular beam placed at its nominal focus (sources.py:299-345)."""
    x = xmin + np.random.rand(num) * (xmax - xmin)
    y = ymin + np.random.rand(num) * (ymax - ymin)
    rho = np.sqrt(x ** 2 + y ** 2)
    theta = np.arctan2(y, x)
    return _converging(x, y, rho, theta, zset, num, lscat)


@_uploaded
def rectbeam(xhalfwidth, yhalfwidth, num):
    """Uniform rectangular beam in +z (sources.py:348-379)."""
    x = (np.random.rand(num) - .5) * 2 * xhalfwidth
    y = (np.random.rand(num) - .5) * 2 * yhalfwidth
    zero = np.repeat(0., num)
    return _bundle(zero, x, y, zero, zero, zero, np.repeat(1., num), zero, zero, zero)


@_uploaded
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


@_uploaded
def fanBeam(xang, yang, num):
    """Rectangular fan of rays from a point (sources.py:418-442)."""
    xa, ya = np.meshgrid(np.linspace(-xang, xang, num), np.linspace(-yang, yang, num))
    return _fan(xa.flatten(), ya.flatten())


@_uploaded
def circFan(halfang, rings, arms):
    """Circular fan of rays from a point: ``rings`` radii x ``arms`` azimuths (sources.py:444-471)."""
    rad = np.linspace(0, halfang, rings)
    az = np.linspace(0, 2 * np.pi, arms + 1)[0:-1]
    rr, aa = np.meshgrid(rad, az)
    xx = np.sin(rr) * np.cos(aa)
    yy = np.sin(rr) * np.sin(aa)
    return _fan(xx.flatten(), yy.flatten())
