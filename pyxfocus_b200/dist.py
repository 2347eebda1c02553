"""Sharded analyses: one process per GPU, each holding a contiguous slice of the bundle.

Tracing needs no communication (rays are independent).  Only the reductions do
(SURVEY.md 8e):

* centroid / rms / image plane: all-reduce(sum) of <= 9 doubles;
* unweighted HPD: distributed exact radix select -- per pass every rank histograms one
  13-bit digit of the radii that still match the resolved prefix, the 2x8192 uint64
  histogram (128 KiB) is all-reduced, and every rank narrows the prefix identically.
  Five passes resolve the full 64-bit pattern, i.e. the exact two middle order statistics
  of the GLOBAL bundle (np.median semantics).

* weighted HPD (analyses.py:88-97 with weights): every rank sorts its own radii (the library's
  radix sort) and prefix-sums its weights in that order; the global weighted quantile is then a
  64-ary search over the 64-bit key space -- per step each rank looks up "my weight at or below this
  key" for 63 pivots per quantile in its sorted shard and one [2, 63] all-reduce merges them -- i.e. a
  merge of the per-rank sorted runs that never moves a ray (11 latency-bound all-reduces).  Large bundles
  do not even sort the shard: a gathered sample brackets the two crossings, one pass collects the ~1 % of
  (radius, weight) pairs inside, and only those windows are sorted and merged (``hpd_weighted_bracketed``).

* full sorted CDF (``rhocdf``): a sample sort -- local sort, splitters agreed on from all-gathered local
  quantiles, one all-to-all of (radius, weight) pairs, stable merge, prefix sums with all-gathered offsets.

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
    """libpxf select primitives on one shard (device-resident state).  The gloo tests replace
    this class by a numpy stand-in with the same interface to exercise the drivers below."""

    def __init__(self, x, y, cxy):
        self.x, self.y, self.cxy = x, y, cxy
        self.dev = x.device
        self.L = _lib.lib()
        self.s = stream_ptr(self.dev)
        self.num = int(x.shape[0])
        self.state = self._new_state()
        self.hist, self.nan = self._views(self.state)
        self.keys = None           # key buffer the histogram passes read (None: radii of x,y)
        self.key_count = None

    # -- state helpers
    def _new_state(self):
        nbytes = int(self.L.pxf_select_state_bytes())
        return torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=self.dev)   # 8-byte aligned

    def _views(self, state):
        base = state.data_ptr()
        ho = (int(self.L.pxf_select_hist_ptr(base)) - base) // 8
        no = (int(self.L.pxf_select_nan_ptr(base)) - base) // 8
        return state[ho:ho + 2 * 8192], state[no:no + 1]

    def schedule(self):
        shift, bits = ctypes.c_int32(), ctypes.c_int32()
        n = self.L.pxf_select_schedule(0, ctypes.byref(shift), ctypes.byref(bits))
        out = []
        for p in range(n):
            self.L.pxf_select_schedule(p, ctypes.byref(shift), ctypes.byref(bits))
            out.append((shift.value, bits.value))
        return out

    # -- five-pass select over the shard's radii (or over a key buffer)
    def begin(self, k0, k1):
        _lib.check(self.L.pxf_select_begin(self.state.data_ptr(), k0, k1, self.s))

    def histogram(self, shift, bits):
        """Add this shard's digit histogram; returns the tensor to all-reduce."""
        if self.keys is not None:
            _lib.check(self.L.pxf_select_hist_keys(self.keys.data_ptr(), self.keys.shape[0],
                                                   self.key_count.data_ptr() if self.key_count is not None else None,
                                                   shift, bits, self.state.data_ptr(), self.s))
        elif self.num > 0:
            _lib.check(self.L.pxf_select_hist(self.x.data_ptr(), self.y.data_ptr(), None, self.num,
                                              self.cxy.data_ptr(), shift, bits, self.state.data_ptr(), self.s))
        return self.hist[:2 << bits]

    def narrow(self, bits):
        _lib.check(self.L.pxf_select_narrow(bits, self.state.data_ptr(), self.s))

    def nan_count(self):
        return self.nan

    def finish(self, total, read=True):
        """(2*median, lower, upper, valid); with read=False the 4 doubles stay on the device
        in ``self.last`` and nothing is returned (no host sync)."""
        out = torch.empty(4, dtype=torch.float64, device=self.dev)
        _lib.check(self.L.pxf_select_finish(self.state.data_ptr(), total, out.data_ptr(), self.s))
        self.last = out
        if not read:
            return None
        h = out.cpu().numpy()
        return float(h[0]), float(h[1]), float(h[2]), bool(h[3] != 0.)

    # -- bracketed select
    def bracket_params(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        ns = int(self.L.pxf_bracket_samples())
        self.L.pxf_bracket_sample_ranks(ns, ctypes.byref(a), ctypes.byref(b))
        return int(self.L.pxf_bracket_min_num()), ns, int(a.value), int(b.value)

    def sample(self, nsamp):
        """nsamp strided radii of this shard (device tensor)."""
        keys = torch.empty(nsamp, dtype=torch.float64, device=self.dev)
        _lib.check(self.L.pxf_select_sample(self.x.data_ptr(), self.y.data_ptr(), self.num, self.cxy.data_ptr(),
                                            nsamp, keys.data_ptr(), self.s))
        return keys

    def use_keys(self, keys, count=None):
        self.keys, self.key_count = keys, count

    def collect(self, lohi):
        """One pass over the shard: returns (candidate buffer, local counters[4] int64)."""
        cap = int(self.L.pxf_bracket_capacity(self.num))
        cand = torch.empty(cap, dtype=torch.float64, device=self.dev)
        counters = torch.zeros(5, dtype=torch.int64, device=self.dev)
        _lib.check(self.L.pxf_bracket_collect(self.x.data_ptr(), self.y.data_ptr(), self.num, self.cxy.data_ptr(),
                                              lohi.data_ptr(), cand.data_ptr(), cap, counters.data_ptr(), self.s))
        # (fill_ takes the scalar as a kernel argument; `counters[4] = cap` would stage it through a
        # pageable host tensor, i.e. block the host until the stream -- the trace kernel -- has drained)
        counters[3:4].copy_((counters[1:2] > cap).to(torch.int64))   # this shard's buffer overflowed
        counters[4:5].fill_(cap)                                      # capacities are summed by the all-reduce
        return cand, counters

    # -- fused small selects (single-CTA kernels; the collectives go between them)
    def small_select(self, keys, seg_counts, nseg, seg_cap, ra, rb, npass, use_scan=False, read=False):
        """Order statistics ra, rb of nseg x seg_cap keys -> device [a+b, a, b, valid]."""
        out = torch.empty(4, dtype=torch.float64, device=self.dev)
        _lib.check(self.L.pxf_small_select(keys.data_ptr(), seg_counts.data_ptr() if seg_counts is not None else None,
                                           nseg, seg_cap, ra, rb, npass, self.fs.data_ptr() if use_scan else None,
                                           out.data_ptr(), self.s))
        self.last = out
        if not read:
            return out
        h = out.cpu().numpy()
        return float(h[0]), float(h[1]), float(h[2]), bool(h[3] != 0.)

    def cand_hist(self, cand, count, lohi):
        """This shard's candidates in linear bins over the bracket (int32 tensor to all-reduce)."""
        fh = torch.zeros(int(self.L.pxf_fast_nbins()), dtype=torch.int32, device=self.dev)
        _lib.check(self.L.pxf_cand_hist(cand.data_ptr(), cand.shape[0], count.data_ptr(), lohi.data_ptr(),
                                        fh.data_ptr(), self.s))
        return fh

    def cand_scan(self, fhist, counters, k0, k1):
        self.fs = torch.zeros((int(self.L.pxf_fastsel_bytes()) + 7) // 8, dtype=torch.int64, device=self.dev)
        _lib.check(self.L.pxf_cand_scan(fhist.data_ptr(), counters.data_ptr(), k0, k1, self.fs.data_ptr(), self.s))

    def cand_gather(self, cand, count, lohi):
        """(fin buffer, its int32 count) of this shard: the candidates of the chosen bins."""
        cap = int(self.L.pxf_fast_fincap())
        fin = torch.empty(cap, dtype=torch.float64, device=self.dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.dev)
        _lib.check(self.L.pxf_cand_gather(cand.data_ptr(), cand.shape[0], count.data_ptr(), lohi.data_ptr(),
                                          self.fs.data_ptr(), fin.data_ptr(), cnt.data_ptr(), self.s))
        return fin, cnt

    def begin_bracket(self, k0, k1, counters):
        """counters: the all-reduced [below, inside, nan, overflowed, capacity] totals."""
        _lib.check(self.L.pxf_select_begin_bracket(self.state.data_ptr(), k0, k1, counters.data_ptr(), -1, self.s))


def _five_passes(sel, group, reduce=True):
    for shift, bits in sel.schedule():
        h = sel.histogram(shift, bits)
        if reduce:
            all_reduce_sum(h, group)
        sel.narrow(bits)


def select_median_pair(sel, total, group=None):
    """Distributed exact select of the two middle order statistics of ``total`` keys spread
    over the ranks of ``group``: five passes, each all-reducing a 2x8192 histogram.
    Returns (2*median, lower middle, upper middle, valid), identical on every rank."""
    k0 = (total - 1) // 2 if total > 0 else 0
    k1 = total // 2 if total > 0 else 0
    sel.use_keys(None)
    sel.begin(k0, k1)
    if total > 0:
        _five_passes(sel, group)
    all_reduce_sum(sel.nan_count(), group)
    return sel.finish(total)


def bracket_median_pair(sel, total, min_shard, group=None, allk=None, read=True):
    """The same statistic with ONE pass over the shards instead of five (see include/pxf.h,
    "Bracketed select") and four small collectives instead of ten 128 KiB all-reduces:

      strided sample of every shard -> ALL-GATHER -> one-CTA select of two sample order statistics
      = the identical bracket [lo,hi] on every rank -> one pass over the shard (count below, collect
      inside) -> ALL-REDUCE the 5 counters -> candidates into 4096 linear bins over the bracket ->
      ALL-REDUCE the bins -> scan (which bins hold the two middle ranks) -> each rank's few hundred
      keys of those bins -> ALL-GATHER -> one-CTA exact select of the union.

    Returns None when the bundle is too small for a bracket (caller uses ``select_median_pair``); a
    result with valid=False means the bracket missed or a buffer overflowed (ties) and the caller
    must fall back as well."""
    world = _world(group)
    min_num, nsamp, ra, rb = sel.bracket_params()
    per = nsamp // world
    if total < min_num or min_shard < max(4 * per, 1):
        return None
    if allk is None:                 # (hpd() below gathers the sample together with the centroid sums)
        mine = sel.sample(per)
        if world > 1:
            allk = torch.empty(per * world, dtype=mine.dtype, device=mine.device)
            td.all_gather_into_tensor(allk, mine, group=group)
        else:
            allk = mine
    n_s = per * world
    scale = n_s / float(nsamp)
    lohi = sel.small_select(allk, None, 1, n_s, max(0, int(ra * scale)), min(n_s - 1, int(rb * scale) + 1), 3)
    cand, local = sel.collect(lohi)
    fh = sel.cand_hist(cand, local[1:2], lohi)
    if world > 1:
        # counters and bins travel in ONE all-reduce, the per-rank key lists and their lengths in ONE
        # all-gather: every collective here is latency bound
        both = torch.cat([local, fh.to(torch.int64)])
        all_reduce_sum(both, group)
        glob, fh = both[:local.shape[0]].contiguous(), both[local.shape[0]:].to(torch.int32)
    else:
        glob = local
    sel.cand_scan(fh, glob, (total - 1) // 2, total // 2)
    fin, cnt = sel.cand_gather(cand, local[1:2], lohi)
    cap = fin.shape[0]
    if world > 1:
        mine2 = torch.cat([fin, cnt.to(fin.dtype)])
        gathered = torch.empty((cap + 1) * world, dtype=fin.dtype, device=fin.device)
        td.all_gather_into_tensor(gathered, mine2, group=group)
        gathered = gathered.view(world, cap + 1)
        fin_all = gathered[:, :cap].contiguous()
        cnt_all = gathered[:, cap].to(torch.int32).contiguous()
    else:
        fin_all, cnt_all = fin, cnt
    return sel.small_select(fin_all, cnt_all, world, cap, 0, 0, 5, use_scan=True, read=read)


def hpd(rays, group=None, return_stats=False, sums=None, total=None, min_shard=None, weights=None, out=None):
    """``weights`` given: the weighted statistic (``hpd_weighted``).  Otherwise: unweighted HPD (2 x median radius about the global centroid) of a sharded bundle;
    exact, identical on every rank.  ``sums``: this shard's centroid sums from
    ``Program.run(..., sums=...)`` (saves the local pass that computes them).  ``total`` /
    ``min_shard``: global ray count and smallest shard size when the caller knows them (e.g.
    equal shards) -- then nothing is read back to the host before the final result, so the
    whole analysis is enqueued behind the trace kernel without a sync.  ``out`` (device float64[4],
    needs ``total``/``min_shard`` and a bundle large enough for the bracketed select): the result
    [HPD, lower, upper, valid] is left there and NOTHING is read back -- for pipelines that analyse
    bundle after bundle and read the numbers at the end; valid == 0 means "repeat without out"."""
    if weights is not None:
        return hpd_weighted(rays, weights, group)
    flush(rays)
    x, y = rays[1:3]
    dev = x.device
    L = _lib.lib()
    world = _world(group)
    sums = _sums(0, rays, None, 0., 0.) if sums is None else sums.clone()
    with torch.cuda.device(dev):
        cxy = torch.empty(2, dtype=torch.float64, device=dev)
        sel = CudaSelect(x, y, cxy)
        min_num, nsamp, _, _ = sel.bracket_params()
        per = nsamp // world
        res = None
        if total is not None and min_shard is not None and total >= min_num and min_shard >= max(4 * per, 1):
            # ONE all-gather carries every rank's centroid sums and its strided (x,y) sample; each rank
            # then adds the sums in rank order (same bits everywhere) and takes the sample radii about
            # the global centroid.
            pack = torch.empty(4 + 2 * per, dtype=torch.float64, device=dev)
            _lib.check(L.pxf_sample_pack(x.data_ptr(), y.data_ptr(), x.shape[0], sums.data_ptr(), per, pack.data_ptr(),
                                         stream_ptr(dev)))
            if world > 1:
                gathered = torch.empty(pack.shape[0] * world, dtype=torch.float64, device=dev)
                td.all_gather_into_tensor(gathered, pack, group=group)
            else:
                gathered = pack
            allk = torch.empty(per * world, dtype=torch.float64, device=dev)
            _lib.check(L.pxf_sample_radii(gathered.data_ptr(), world, per, sums.data_ptr(), cxy.data_ptr(),
                                          allk.data_ptr(), stream_ptr(dev)))
            res = bracket_median_pair(sel, total, min_shard, group, allk=allk, read=out is None)
            if out is not None:
                out.copy_(sel.last)
                return out
        else:
            if out is not None:
                raise ValueError("dist.hpd(out=...) needs total/min_shard and a bundle large enough for the bracket")
            all_reduce_sum(sums[:4], group)
            if total is None or min_shard is None:
                cnt = torch.tensor([float(x.shape[0])], dtype=torch.float64, device=dev)
                if world > 1:
                    td.all_reduce(cnt, op=td.ReduceOp.MIN, group=group)
                total = int(round(float(sums[3].item())))
                min_shard = int(round(float(cnt.item())))
            _lib.check(L.pxf_centroid_from_sums(sums.data_ptr(), cxy.data_ptr(), stream_ptr(dev)))
            res = bracket_median_pair(sel, total, min_shard, group)
        if res is None or not res[3]:
            res = select_median_pair(sel, total, group)
    return res[:3] if return_stats else res[0]


# ------------------------------------------------------------------ weighted HPD
class CudaWeighted:
    """Local half of the weighted quantile: radii about the global centroid, sorted, with the prefix
    sums of the weights in sorted order (libpxf kernels: pxf_rho, pxf_argsort, pxf_cumsum_gather).
    The gloo tests replace it with a CPU stand-in of the same interface."""

    def __init__(self, rays, weights, cx, cy):
        from . import analyses
        x, y = rays[1:3]
        dev = x.device
        L = _lib.lib()
        num = x.shape[0]
        w = torch.as_tensor(weights, dtype=torch.float64, device=dev).contiguous()
        r = torch.empty_like(x)
        self.cum = torch.empty_like(x)
        with torch.cuda.device(dev):
            if num > 0:
                _lib.check(L.pxf_rho(x.data_ptr(), y.data_ptr(), num, cx, cy, r.data_ptr(), stream_ptr(dev)))
                rs, idx = analyses.argsort(r)
                scratch = torch.empty(int(L.pxf_scan_scratch_bytes(num)), dtype=torch.uint8, device=dev)
                _lib.check(L.pxf_cumsum_gather(w.data_ptr(), idx.data_ptr(), num, self.cum.data_ptr(),
                                               scratch.data_ptr(), stream_ptr(dev)))
            else:
                rs = r
        self.keys = rs.view(torch.int64)          # radii are >= 0: the bit pattern orders like the value
        self.device = dev


def _weight_at_or_below(keys, cum, key, strict=False):
    """Weight of this shard's sorted (keys, cum) run at or below each entry of ``key`` (below when strict)."""
    if keys.shape[0] == 0:
        return torch.zeros(key.shape, dtype=torch.float64, device=key.device)
    cnt = torch.searchsorted(keys, key.reshape(-1), right=not strict).reshape(key.shape)
    val = cum[torch.clamp(cnt - 1, min=0)]
    return torch.where(cnt > 0, val, torch.zeros_like(val))


def _largest_key_below(keys, key):
    """Largest key of this shard's sorted run that is < key (0-d), or -1."""
    if keys.shape[0] == 0:
        return torch.full((), -1, dtype=torch.int64, device=key.device)
    cnt = torch.searchsorted(keys, key.reshape(1), right=False)[0]
    val = keys[torch.clamp(cnt - 1, min=0)]
    return torch.where(cnt > 0, val, torch.full_like(val, -1))


_KARY = 63          # pivots per step: 11 steps of one [2, 63] all-reduce resolve the 63-bit key space
_KARY_STEPS = 11


def weighted_quantile_radii(runs, qs, total_weight, group=None, offsets=None, key_ranges=None):
    """For each quantile q = qs[i]: the radius r[argmin |cdf - q|] of the GLOBAL sorted (radius, weight) sequence
    formed by every rank's sorted run ``runs[i] = (keys, cum)`` (cdf = (offsets[i] + cumulative weight) / total;
    first minimum, as numpy's argmin): the smallest key whose global cumulative weight reaches q*W, or its
    predecessor when that one is at least as close.  A (K+1)-ary search over the 64-bit key space: per step
    every rank looks up "my weight at or below this key" for K pivots per quantile in its sorted run and ONE
    [len(qs), K] all-reduce merges them -- the rays never move.  Everything stays on the device and is
    identical on every rank.  Returns (radii [len(qs)] float64, valid [len(qs)] bool): valid is False when no key
    of the runs reaches q, or when the predecessor would lie outside the runs although weight lies below them
    (``offsets[i]`` > 0) -- the bracketed caller then falls back to the full sort."""
    dev = runs[0][0].device
    nq = len(qs)
    if dev.type == "cuda" and nq == 2:
        return _weighted_quantile_radii_cuda(runs, qs, total_weight, group, offsets, key_ranges)
    q = torch.tensor(list(qs), dtype=torch.float64, device=dev)
    off = torch.zeros(nq, dtype=torch.float64, device=dev) if offsets is None else offsets.to(torch.float64)
    lo = torch.zeros(nq, dtype=torch.int64, device=dev)                              # invariant: answer in [lo, hi]
    hi = torch.full((nq,), 0x7ff0000000000000, dtype=torch.int64, device=dev)        # +Inf pattern
    if key_ranges is not None:
        # every key of run i lies in key_ranges[i] (identical on every rank): a narrower space, fewer steps
        lo = torch.tensor([max(int(a), 0) for a, _ in key_ranges], dtype=torch.int64, device=dev)
        hi = torch.tensor([min(int(b), 0x7ff0000000000000) for _, b in key_ranges], dtype=torch.int64, device=dev)
    steps = _merge_steps(key_ranges)[1] if key_ranges is not None and nq == 2 else _KARY_STEPS
    jj = torch.arange(1, _KARY + 1, dtype=torch.int64, device=dev)                   # j+1

    def cdf_at(keys2d, strict=False):
        loc = torch.stack([_weight_at_or_below(runs[i][0], runs[i][1], keys2d[i], strict) for i in range(nq)])
        return (all_reduce_sum(loc, group) + off.reshape(nq, 1)) / total_weight

    for _ in range(steps):
        n = hi - lo + 1
        step, rem = n // (_KARY + 1), n % (_KARY + 1)
        piv = lo.reshape(nq, 1) + jj * step.reshape(nq, 1) + torch.minimum(jj.expand(nq, _KARY), rem.reshape(nq, 1)) - 1
        piv = torch.minimum(piv, hi.reshape(nq, 1))
        ge = cdf_at(piv) >= q.reshape(nq, 1)
        anyge = ge.any(dim=1)
        first = torch.argmax(ge.to(torch.int8), dim=1)                               # first pivot reaching q
        pj = torch.gather(piv, 1, first.reshape(nq, 1)).reshape(nq)
        pprev = torch.gather(piv, 1, torch.clamp(first - 1, min=0).reshape(nq, 1)).reshape(nq)
        new_hi = torch.where(anyge, pj, hi)
        new_lo = torch.where(anyge, torch.where(first > 0, pprev + 1, lo), piv[:, -1] + 1)
        lo, hi = torch.minimum(new_lo, new_hi), new_hi
    k_hi = hi
    c_hi = cdf_at(k_hi.reshape(nq, 1)).reshape(nq)
    k_lo = torch.stack([_largest_key_below(runs[i][0], k_hi[i]) for i in range(nq)])
    if _world(group) > 1:
        td.all_reduce(k_lo, op=td.ReduceOp.MAX, group=group)
    c_lo = cdf_at(k_hi.reshape(nq, 1), strict=True).reshape(nq)
    take_lo = (k_lo >= 0) & (torch.abs(c_lo - q) <= torch.abs(c_hi - q))
    key = torch.where(take_lo, k_lo, k_hi)
    valid = (c_hi >= q) & ((k_lo >= 0) | (off == 0.))
    return key.view(torch.float64), valid


def _merge_steps(key_ranges):
    if key_ranges is None:
        return (0, 0x7ff0000000000000, 0, 0x7ff0000000000000), _KARY_STEPS
    (a0, b0), (a1, b1) = key_ranges
    a0, a1 = max(int(a0), 0), max(int(a1), 0)
    b0, b1 = max(min(int(b0), 0x7ff0000000000000), a0), max(min(int(b1), 0x7ff0000000000000), a1)
    width = max(b0 - a0 + 1, b1 - a1 + 1, 1)
    steps = 1
    while (_KARY + 1) ** steps < width:
        steps += 1
    return (a0, b0, a1, b1), steps


def _weighted_quantile_radii_cuda(runs, qs, total_weight, group, offsets, key_ranges):
    """``weighted_quantile_radii`` with libpxf kernels (pxf_wq_merge_*): two launches and one all-reduce per step
    instead of ~25 small tensor ops.  Same pivots, same narrowing, same result."""
    dev = runs[0][0].device
    L = _lib.lib()
    (k0, c0), (k1, c1) = runs
    k0, c0, k1, c1 = k0.contiguous(), c0.contiguous(), k1.contiguous(), c1.contiguous()
    n0, n1 = int(k0.shape[0]), int(k1.shape[0])
    W = total_weight.reshape(1).to(torch.float64).contiguous()
    off = None if offsets is None else offsets.to(torch.float64).contiguous()
    (a0, b0, a1, b1), steps = _merge_steps(key_ranges)
    world = _world(group)
    with torch.cuda.device(dev):
        st = torch.zeros(int(L.pxf_wq_merge_state_bytes()) // 8, dtype=torch.float64, device=dev)
        probe = torch.empty(2 * _KARY, dtype=torch.float64, device=dev)
        fin = torch.empty(4, dtype=torch.float64, device=dev)
        s = stream_ptr(dev)
        offp = off.data_ptr() if off is not None else None
        _lib.check(L.pxf_wq_merge_begin(st.data_ptr(), a0, b0, a1, b1, s))
        for _ in range(steps):
            _lib.check(L.pxf_wq_merge_probe(k0.data_ptr(), c0.data_ptr(), n0, k1.data_ptr(), c1.data_ptr(), n1,
                                            st.data_ptr(), _KARY, probe.data_ptr(), s))
            all_reduce_sum(probe, group)
            _lib.check(L.pxf_wq_merge_narrow(st.data_ptr(), probe.data_ptr(), offp, W.data_ptr(), float(qs[0]), float(qs[1]),
                                             _KARY, s))
        _lib.check(L.pxf_wq_merge_final_probe(k0.data_ptr(), c0.data_ptr(), n0, k1.data_ptr(), c1.data_ptr(), n1,
                                              st.data_ptr(), fin.data_ptr(), s))
        all_reduce_sum(fin, group)
        if world > 1:
            td.all_reduce(st[4:6].view(torch.int64), op=td.ReduceOp.MAX, group=group)      # klo[2]
        _lib.check(L.pxf_wq_merge_finish(st.data_ptr(), fin.data_ptr(), offp, W.data_ptr(), float(qs[0]), float(qs[1]), s))
    return st[6:8].clone(), st[8:10] != 0.


def weighted_quantile_radius(loc, q, total_weight, group=None):
    """One quantile of the full sorted runs (``loc.keys``, ``loc.cum``); see ``weighted_quantile_radii``."""
    r, _ = weighted_quantile_radii([(loc.keys, loc.cum)], [q], total_weight, group)
    return r[0]


class CudaWeightedBracket:
    """Local half of the bracketed weighted quantile (libpxf pxf_wq_* kernels, pxf_wquant.cu): strided
    (radius, weight) sample, brackets from the gathered sample, one collect pass, sorted candidate windows.
    The gloo tests replace it with a numpy stand-in of the same interface."""

    def __init__(self, rays, weights, cx, cy):
        x, y = rays[1:3]
        self.x, self.y = x, y
        self.device = x.device
        self.num = x.shape[0]
        self.w = torch.as_tensor(weights, dtype=torch.float64, device=x.device).contiguous()
        self.cxy = torch.tensor([cx, cy], dtype=torch.float64, device=x.device)
        self.L = _lib.lib()
        self.state = torch.zeros(int(self.L.pxf_wq_state_bytes()) // 8, dtype=torch.float64, device=x.device)

    def params(self, total):
        return int(self.L.pxf_wq_min_num()), int(self.L.pxf_wq_samples(int(total)))

    def sample(self, nsamp):
        """[2, nsamp]: radii and weights of a strided sample, weights scaled by rays-per-sample so that samples
        of differently sized shards combine into one unbiased weighted cdf; padded with (+Inf, 0)."""
        out = torch.empty(2, nsamp, dtype=torch.float64, device=self.device)
        out[0].fill_(float("inf"))
        out[1].zero_()
        take = min(nsamp, self.num)
        if take > 0:
            with torch.cuda.device(self.device):
                _lib.check(self.L.pxf_wq_sample(self.x.data_ptr(), self.y.data_ptr(), self.w.data_ptr(), self.num,
                                                self.cxy.data_ptr(), take, out[0].data_ptr(), out[1].data_ptr(),
                                                stream_ptr(self.device)))
            out[1, :take].mul_(self.num / take)
        return out

    def _sorted(self, r, w, digits):
        """(keys sorted on the bytes in ``digits``, prefix sums of the weights in that order); no read-back."""
        n = r.shape[0]
        rs = torch.empty_like(r)
        idx = torch.empty(n, dtype=torch.int64, device=self.device)
        cum = torch.empty_like(r)
        with torch.cuda.device(self.device):
            sscr = torch.empty(int(self.L.pxf_sort_scratch_bytes(n)), dtype=torch.uint8, device=self.device)
            _lib.check(self.L.pxf_argsort_digits(r.data_ptr(), n, rs.data_ptr(), idx.data_ptr(), sscr.data_ptr(), digits,
                                                 stream_ptr(self.device)))
            scratch = torch.empty(int(self.L.pxf_scan_scratch_bytes(n)), dtype=torch.uint8, device=self.device)
            _lib.check(self.L.pxf_cumsum_gather(w.data_ptr(), idx.data_ptr(), n, cum.data_ptr(), scratch.data_ptr(),
                                                stream_ptr(self.device)))
        return rs, cum

    def set_brackets(self, gathered):
        """gathered: [world, 2, nsamp] samples of every rank (identical everywhere) -> brackets in the state."""
        r = gathered[:, 0, :].reshape(-1).contiguous()
        w = gathered[:, 1, :].reshape(-1).contiguous()
        # a bracket only needs the sample ordered on the top 32 bits of the radius pattern (pxf_wquant.cu)
        rs, cum = self._sorted(r, w, 0xF0)
        with torch.cuda.device(self.device):
            _lib.check(self.L.pxf_wq_brackets(rs.data_ptr(), cum.data_ptr(), rs.shape[0], 32, self.state.data_ptr(),
                                              stream_ptr(self.device)))

    def collect(self, cap):
        """One pass over the shard.  Returns float64[5]: weight below bracket 0 / 1, candidates in bracket
        0 / 1 beyond the capacity (0 = fits), rays the bracketed path refuses (NaN radius, NaN/negative weight)."""
        dev = self.device
        self.cand = torch.empty(4, max(cap, 1), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            scratch = torch.empty(int(self.L.pxf_wq_collect_scratch_bytes()), dtype=torch.uint8, device=dev)
            _lib.check(self.L.pxf_wq_collect(self.x.data_ptr(), self.y.data_ptr(), self.w.data_ptr(), self.num,
                                             self.cxy.data_ptr(), self.state.data_ptr(), self.cand[0].data_ptr(),
                                             self.cand[1].data_ptr(), self.cand[2].data_ptr(), self.cand[3].data_ptr(),
                                             cap, scratch.data_ptr(), stream_ptr(dev)))
        cnt = self.state[6:9].view(torch.int64)                   # count[0], count[1], nbad
        self.counts = cnt[:2]
        self.cap = cap
        over = torch.clamp(cnt[:2] - cap, min=0).to(torch.float64)
        return torch.cat([self.state[4:6], over, cnt[2:3].to(torch.float64)])

    def windows(self):
        """[(keys, cum)] x 2: this shard's candidates of each bracket, sorted, with prefix weights."""
        host = self.state[:9].cpu()                               # one read-back: brackets and counts
        n0, n1 = (int(v) for v in host[6:8].view(torch.int64))
        pat = host[:4].view(torch.int64)
        self.key_range = [(int(pat[0]), int(pat[1])), (int(pat[2]), int(pat[3]))]
        out = []
        for b, n in ((0, n0), (1, n1)):
            n = min(n, self.cap)
            diff = self.key_range[b][0] ^ self.key_range[b][1]
            digits = sum(1 << d for d in range(8) if (diff >> (8 * d)) != 0) or 1
            if n == 0:
                out.append((torch.empty(0, dtype=torch.int64, device=self.device),
                            torch.empty(0, dtype=torch.float64, device=self.device)))
                continue
            rs, cum = self._sorted(self.cand[2 * b, :n].contiguous(), self.cand[2 * b + 1, :n].contiguous(), digits)
            out.append((rs.view(torch.int64), cum))
        return out


def hpd_weighted_bracketed(loc, total, total_weight, group=None):
    """Bracketed weighted HPD of a sharded bundle (the multi-GPU form of pxf_hpd_weighted_bracket): returns
    (hpd 0-d tensor, valid bool).  Collectives: one all-gather of the samples, one all-reduce of 5 doubles, then
    the (K+1)-ary merge of the sorted candidate windows (11 all-reduces of [2, 63] doubles + 3 small ones)."""
    world = _world(group)
    dev = loc.device
    _, nsamp = loc.params(total)
    per = max(nsamp // world, 1024)
    mine = loc.sample(per)
    if world > 1:
        flat = torch.empty(world * mine.numel(), dtype=torch.float64, device=dev)
        td.all_gather_into_tensor(flat, mine.reshape(-1).contiguous(), group=group)
        gathered = flat.reshape((world,) + tuple(mine.shape))
    else:
        gathered = mine.reshape((1,) + tuple(mine.shape))
    loc.set_brackets(gathered)
    cap = max(int(loc.num) // 8, 65536)
    st = all_reduce_sum(loc.collect(cap), group)
    below, bad = st[:2], st[2:]
    runs = loc.windows()
    r, valid = weighted_quantile_radii(runs, [.25, .75], total_weight, group, offsets=below,
                                       key_ranges=getattr(loc, "key_range", None))
    ok = bool(valid.all().item()) and float(bad.sum().item()) == 0.
    return r[1] - r[0], ok


def hpd_weighted(rays, weights, group=None, local_cls=CudaWeighted, bracket_cls=CudaWeightedBracket):
    """Weighted HPD of a sharded bundle: r[argmin|cdf-.75|] - r[argmin|cdf-.25|] about the global
    weighted centroid (analyses.py:88-97), identical on every rank.  Large bundles take the bracketed path
    (no rank sorts more than the ~1 % of its rays next to the two crossings); small ones, or a bracket miss,
    the full local sort."""
    flush(rays)
    dev = rays[1].device
    if _world(group) == 1 and local_cls is CudaWeighted and bracket_cls is CudaWeightedBracket:
        from . import analyses
        return analyses.hpd(rays, weights=weights)
    s = all_reduce_sum(torch.cat([_sums(0, rays, weights, 0., 0.)[:3],
                                  torch.tensor([float(rays[1].shape[0])], dtype=torch.float64, device=dev)]), group)
    h = s.cpu().numpy()
    cx, cy, total = float(h[1] / h[0]), float(h[2] / h[0]), int(round(h[3]))
    if bracket_cls is not None:
        br = bracket_cls(rays, weights, cx, cy)
        if total >= br.params(total)[0]:
            res, ok = hpd_weighted_bracketed(br, total, s[0], group)
            if ok:
                return float(res)
    loc = local_cls(rays, weights, cx, cy)
    n = loc.keys.shape[0]
    wl = loc.cum[n - 1].clone() if n > 0 else torch.zeros((), dtype=torch.float64, device=loc.device)
    W = all_reduce_sum(wl, group)
    r, _ = weighted_quantile_radii([(loc.keys, loc.cum)] * 2, [.25, .75], W, group)
    return float(r[1] - r[0])


# ------------------------------------------------------------------ sharded rhocdf: sample sort
class CudaSortedRun:
    """Local half of the sharded ``rhocdf``: this shard's radii about the global centroid, sorted, with their
    weights in the same order (libpxf kernels: pxf_rho, pxf_argsort).  The gloo tests use a numpy stand-in."""

    def __init__(self, rays, weights, cx, cy):
        from . import analyses
        x, y = rays[1:3]
        dev = x.device
        num = x.shape[0]
        self.device = dev
        w = None if weights is None else torch.as_tensor(weights, dtype=torch.float64, device=dev).contiguous()
        if num == 0:
            self.r = torch.empty(0, dtype=torch.float64, device=dev)
            self.w = torch.empty(0, dtype=torch.float64, device=dev)
            return
        r = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pxf_rho(x.data_ptr(), y.data_ptr(), num, cx, cy, r.data_ptr(), stream_ptr(dev)))
        self.r, idx = analyses.argsort(r)
        self.w = torch.ones_like(r) if w is None else w[idx]

    @staticmethod
    def merge(r, w):
        """Sort the concatenation of the received runs (stable: ties keep source-rank order)."""
        from . import analyses
        if r.shape[0] == 0:
            return r, w
        rs, idx = analyses.argsort(r.contiguous())
        return rs, w[idx]

    @staticmethod
    def prefix(w):
        """Inclusive prefix sums (pxf_cumsum_gather without a permutation)."""
        out = torch.empty_like(w)
        n = w.shape[0]
        if n:
            L = _lib.lib()
            with torch.cuda.device(w.device):
                scratch = torch.empty(int(L.pxf_scan_scratch_bytes(n)), dtype=torch.uint8, device=w.device)
                _lib.check(L.pxf_cumsum_gather(w.contiguous().data_ptr(), None, n, out.data_ptr(), scratch.data_ptr(),
                                               stream_ptr(w.device)))
        return out


_SPLIT_SAMPLES = 1024


def rhocdf(rays, weights=None, cent=True, group=None, local_cls=CudaSortedRun):
    """Sharded ``analyses.rhocdf`` (analyses.py:73-86): the GLOBAL sorted radii and cumulative weights, distributed --
    rank g returns the g-th contiguous slice ``(r, cdf, first)`` of the global sorted order (``first`` = global index
    of its first element); concatenating the slices in rank order gives what the single-GPU call returns.

    A sample sort, the one place on this path with a real exchange step: every rank sorts its shard, the ranks
    all-gather evenly spaced local quantiles and agree on world-1 splitters, each rank cuts its sorted run at the
    splitters and ONE all-to-all moves every (radius, weight) pair to the rank that owns its value range; the
    received runs are merged by a stable sort, prefix-summed, and offset by the all-gathered bucket totals.  The
    cdf is normalised by the total weight (= the reference's cdf.max() for non-negative weights)."""
    flush(rays)
    world = _world(group)
    rank = td.get_rank(group) if world > 1 else 0
    cx, cy = centroid(rays, weights, group) if cent is True else (0., 0.)
    run = local_cls(rays, weights, cx, cy)
    dev = run.device
    n = run.r.shape[0]
    if world == 1:
        cum = local_cls.prefix(run.w)
        return run.r, (cum / cum[-1] if n else cum), 0
    # splitters from evenly spaced local quantiles (+Inf padding for an empty shard)
    if n > 0:
        pos = ((torch.arange(_SPLIT_SAMPLES, device=dev, dtype=torch.float64) + .5) * (n / _SPLIT_SAMPLES)).long().clamp(max=n - 1)
        mine = run.r[pos]
    else:
        mine = torch.full((_SPLIT_SAMPLES,), float("inf"), dtype=torch.float64, device=dev)
    allq = torch.empty(world * _SPLIT_SAMPLES, dtype=torch.float64, device=dev)
    td.all_gather_into_tensor(allq, mine.contiguous(), group=group)
    allq = torch.sort(allq).values
    split = allq[torch.arange(1, world, device=dev) * _SPLIT_SAMPLES]           # world-1 splitters, same on every rank
    # bucket k = (split[k-1], split[k]]: cut the sorted run, exchange counts, then the pairs
    cuts = torch.searchsorted(run.r, split, right=True)
    bounds = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), cuts, torch.full((1,), n, dtype=torch.int64, device=dev)])
    send_counts = bounds[1:] - bounds[:-1]
    recv_counts = torch.empty_like(send_counts)
    td.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    got_r = torch.empty(sum(rc), dtype=torch.float64, device=dev)
    got_w = torch.empty(sum(rc), dtype=torch.float64, device=dev)
    td.all_to_all_single(got_r, run.r.contiguous(), output_split_sizes=rc, input_split_sizes=sc, group=group)
    td.all_to_all_single(got_w, run.w.contiguous(), output_split_sizes=rc, input_split_sizes=sc, group=group)
    r, w = local_cls.merge(got_r, got_w)
    cum = local_cls.prefix(w)
    # offsets: weight and element count of the buckets before mine
    tot = torch.zeros(2 * world, dtype=torch.float64, device=dev)
    mine2 = torch.stack([cum[-1] if r.shape[0] else torch.zeros((), dtype=torch.float64, device=dev),
                         torch.tensor(float(r.shape[0]), dtype=torch.float64, device=dev)])
    td.all_gather_into_tensor(tot, mine2.contiguous(), group=group)
    tot = tot.reshape(world, 2)
    woff = tot[:rank, 0].sum()
    first = int(tot[:rank, 1].sum().item())
    total = tot[:, 0].sum()
    return r, (woff + cum) / total, first
