"""Sharded analyses: one process per GPU, each holding a contiguous slice of the bundle.

Tracing needs no communication (rays are independent).  Only the reductions do
(SURVEY.md 8e):

* centroid / rms / image plane: all-reduce(sum) of <= 9 doubles;
* unweighted HPD: distributed exact radix select -- per pass every rank histograms one
  13-bit digit of the radii that still match the resolved prefix, the 2x8192 uint64
  histogram (128 KiB) is all-reduced, and every rank narrows the prefix identically.
  Five passes resolve the full 64-bit pattern, i.e. the exact two middle order statistics
  of the GLOBAL bundle (np.median semantics).

Collectives go through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the
CPU tests of the host logic).  With ``group=None`` and no initialised process group the
functions degrade to the single-GPU result (world size 1).
"""
import ctypes

import torch
import torch.distributed as td

from . import _lib
from ._call import stream_ptr
from .program import flush


def _world(group):
    if td.is_available() and td.is_initialized():
        return td.get_world_size(group)
    return 1


def all_reduce_sum(t, group=None):
    if _world(group) > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM, group=group)
    return t


def _sums(mode, rays, weights, a, b):
    x, y, z, l, m, n = rays[1:7]
    L = _lib.lib()
    dev = x.device
    out = torch.zeros(16, dtype=torch.float64, device=dev)
    w = None if weights is None else torch.as_tensor(weights, dtype=torch.float64, device=dev).contiguous()
    with torch.cuda.device(dev):
        scratch = torch.empty(int(L.pxf_sums_scratch_bytes()), dtype=torch.uint8, device=dev)
        if x.shape[0] > 0:
            _lib.check(L.pxf_sums(mode, x.data_ptr(), y.data_ptr(), l.data_ptr(), m.data_ptr(), n.data_ptr(),
                                  w.data_ptr() if w is not None else None, x.shape[0], a, b, out.data_ptr(),
                                  scratch.data_ptr(), stream_ptr(dev)))
    return out


def centroid(rays, weights=None, group=None):
    """Global centroid of a sharded bundle."""
    flush(rays)
    s = all_reduce_sum(_sums(0, rays, weights, 0., 0.), group)
    h = s[:3].cpu().numpy()
    return float(h[1] / h[0]), float(h[2] / h[0])


def rmsCentroid(rays, weights=None, group=None):
    """Global RMS radius about the global centroid."""
    cx, cy = centroid(rays, weights, group)
    s = all_reduce_sum(_sums(1, rays, weights, cx, cy), group)
    h = s[:2].cpu().numpy()
    return float((h[1] / h[0]) ** 0.5)


def analyticImagePlane(rays, weights=None, group=None):
    """Global analytic image plane (analyses.py:118-133) from all-reduced sums."""
    flush(rays)
    h = all_reduce_sum(_sums(2, rays, weights, 0., 0.), group)[:9].cpu().numpy()
    W = h[0]
    mx, my, ma, mb = h[1] / W, h[2] / W, h[3] / W, h[4] / W
    bx = h[5] / W - mx * ma
    ax = h[7] / W - ma * ma
    by = h[6] / W - my * mb
    ay = h[8] / W - mb * mb
    return float(-(bx + by) / (ax + ay))


def shard_range(num, rank, world):
    """Contiguous slice [lo, hi) of a num-ray bundle owned by ``rank``:
    concatenating the shards in rank order reproduces the global ray order."""
    per, rem = divmod(int(num), int(world))
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


class CudaSelect:
    """libpxf radix-select primitives on one shard (device-resident state)."""

    def __init__(self, x, y, cxy):
        self.x, self.y, self.cxy = x, y, cxy
        self.dev = x.device
        self.L = _lib.lib()
        self.s = stream_ptr(self.dev)
        nbytes = int(self.L.pxf_select_state_bytes())
        self.state = torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=self.dev)   # 8-byte aligned
        base = self.state.data_ptr()
        ho = (int(self.L.pxf_select_hist_ptr(base)) - base) // 8
        no = (int(self.L.pxf_select_nan_ptr(base)) - base) // 8
        self.hist = self.state[ho:ho + 2 * 8192]
        self.nan = self.state[no:no + 1]

    def schedule(self):
        shift, bits = ctypes.c_int32(), ctypes.c_int32()
        n = self.L.pxf_select_schedule(0, ctypes.byref(shift), ctypes.byref(bits))
        out = []
        for p in range(n):
            self.L.pxf_select_schedule(p, ctypes.byref(shift), ctypes.byref(bits))
            out.append((shift.value, bits.value))
        return out

    def begin(self, k0, k1):
        _lib.check(self.L.pxf_select_begin(self.state.data_ptr(), k0, k1, self.s))

    def histogram(self, shift, bits):
        """Add this shard's digit histogram; returns the tensor to all-reduce."""
        if self.x.shape[0] > 0:
            _lib.check(self.L.pxf_select_hist(self.x.data_ptr(), self.y.data_ptr(), None, self.x.shape[0],
                                              self.cxy.data_ptr(), shift, bits, self.state.data_ptr(), self.s))
        return self.hist[:2 << bits]

    def narrow(self, bits):
        _lib.check(self.L.pxf_select_narrow(bits, self.state.data_ptr(), self.s))

    def nan_count(self):
        return self.nan

    def finish(self, total):
        out = torch.empty(3, dtype=torch.float64, device=self.dev)
        _lib.check(self.L.pxf_select_finish(self.state.data_ptr(), total, out.data_ptr(), self.s))
        h = out.cpu().numpy()
        return float(h[0]), float(h[1]), float(h[2])


def select_median_pair(sel, total, group=None):
    """Drive a distributed exact select of the two middle order statistics of ``total``
    keys spread over the ranks of ``group``.  ``sel`` provides the per-shard primitives
    (``CudaSelect``; the gloo tests substitute a CPU stand-in to exercise this logic).
    Returns (2*median, lower middle, upper middle), identical on every rank."""
    k0 = (total - 1) // 2 if total > 0 else 0
    k1 = total // 2 if total > 0 else 0
    sel.begin(k0, k1)
    if total > 0:
        for shift, bits in sel.schedule():
            h = sel.histogram(shift, bits)
            all_reduce_sum(h, group)
            sel.narrow(bits)
    all_reduce_sum(sel.nan_count(), group)
    return sel.finish(total)


def hpd(rays, group=None, return_stats=False):
    """Unweighted HPD (2 x median radius about the global centroid) of a sharded bundle;
    exact, identical on every rank."""
    flush(rays)
    x, y = rays[1:3]
    dev = x.device
    sums = all_reduce_sum(_sums(0, rays, None, 0., 0.), group)
    total = int(round(float(sums[3].item())))
    with torch.cuda.device(dev):
        cxy = torch.empty(2, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().pxf_centroid_from_sums(sums.data_ptr(), cxy.data_ptr(), stream_ptr(dev)))
        res = select_median_pair(CudaSelect(x, y, cxy), total, group)
    return res if return_stats else res[0]
