"""Multi-rank host logic on CPU: world_size 2 (and 3) over gloo.

The sharded HPD (pyxfocus_b200.dist) is: all-reduce centroid sums -> per pass every rank
histograms one digit of its shard's radii under the resolved prefix -> all-reduce the
histogram -> every rank narrows identically.  The CUDA kernels are replaced here by a numpy
stand-in with the same state machine (prefix / rank / nprefix, digit schedule 13+13+13+13+12),
so the driver ``dist.select_median_pair``, the sharding and the collectives are what is
tested.  The result must equal np.median of the concatenated bundle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

SCHEDULE = [(51, 13), (38, 13), (25, 13), (12, 13), (0, 12)]


class NumpySelect:
    """CPU stand-in for dist.CudaSelect (same interface, same semantics as the kernels)."""

    def __init__(self, x, y, cx, cy, min_num=2000, nsamp=4096):
        self.r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
        self.num = self.r.shape[0]
        self.hist = torch.zeros(2 * 8192, dtype=torch.int64)
        self.nan = torch.zeros(1, dtype=torch.int64)
        self.keys = None
        self.key_count = None
        self.valid = True
        self.min_num, self.nsamp = min_num, nsamp
        self.collect_calls = 0

    def schedule(self):
        return SCHEDULE

    def _cur(self):
        if self.keys is None:
            return self.r
        k = self.keys.numpy()
        if self.key_count is not None:
            k = k[:min(int(self.key_count.item()), k.shape[0])]
        return k

    def use_keys(self, keys, count=None):
        self.keys, self.key_count = keys, count

    def begin(self, k0, k1):
        self.prefix = [0, 0]
        self.rank = [int(k0), int(k1)]
        self.nprefix = 1
        self.valid = True
        self.hist.zero_()
        self.nan.zero_()

    def histogram(self, shift, bits):
        nb = 1 << bits
        r = self._cur()
        isnan = np.isnan(r)
        k = np.ascontiguousarray(r[~isnan]).view(np.uint64)
        if self.keys is None:
            self.nan += int(isnan.sum())
        top = (k >> np.uint64(shift + bits)) if shift + bits < 64 else np.zeros_like(k)
        dig = ((k >> np.uint64(shift)) & np.uint64(nb - 1)).astype(np.int64)
        h = np.zeros(2 * nb, dtype=np.int64)
        sel0 = top == np.uint64(self.prefix[0])
        h[:nb] += np.bincount(dig[sel0], minlength=nb)
        if self.nprefix == 2:
            sel1 = (top == np.uint64(self.prefix[1])) & ~sel0
            h[nb:] += np.bincount(dig[sel1], minlength=nb)
        self.hist[:2 * nb] += torch.from_numpy(h)
        return self.hist[:2 * nb]

    def narrow(self, bits):
        nb = 1 << bits
        h = self.hist.numpy()
        newp = []
        for j in range(2):
            hs = h[nb:2 * nb] if (self.nprefix == 2 and j == 1) else h[:nb]
            c = np.cumsum(hs)
            b = int(np.searchsorted(c, self.rank[j], side="right"))
            base = int(c[b - 1]) if b > 0 else 0
            p = self.prefix[j] if self.nprefix == 2 else self.prefix[0]
            newp.append(((p << bits) | min(b, nb - 1)) & (2 ** 64 - 1))
            self.rank[j] -= base
        self.prefix = newp
        self.nprefix = 1 if newp[0] == newp[1] else 2
        self.hist.zero_()

    def nan_count(self):
        return self.nan

    def finish(self, total, read=True):
        a = np.array([self.prefix[0]], dtype=np.uint64).view(np.float64)[0]
        b = np.array([self.prefix[1]], dtype=np.uint64).view(np.float64)[0]
        med = (a + b) / 2.
        if total == 0 or int(self.nan.item()) > 0:
            med = float("nan")
        self.last = torch.tensor([2. * med, a, b, 1. if self.valid else 0.], dtype=torch.float64)
        if not read:
            return None
        return 2. * med, float(a), float(b), self.valid

    # bracketed select
    def bracket_params(self):
        sigma = .5 / np.sqrt(self.nsamp)
        d = int(np.ceil(6 * sigma * self.nsamp)) + 1
        return self.min_num, self.nsamp, max(0, self.nsamp // 2 - d), min(self.nsamp - 1, self.nsamp // 2 + d)

    def sample(self, nsamp):
        idx = (np.arange(nsamp, dtype=np.int64) * self.num) // nsamp
        return torch.from_numpy(self.r[idx].copy())

    def collect(self, lohi):
        self.collect_calls += 1
        lo, hi = float(lohi[1]), float(lohi[2])
        cap = max(64, self.num // 4)
        r = self.r
        inside = r[(r >= lo) & (r <= hi)]
        cand = torch.zeros(cap, dtype=torch.float64)
        cand[:min(cap, inside.shape[0])] = torch.from_numpy(inside[:cap].copy())
        counters = torch.tensor([int((r < lo).sum()), inside.shape[0], int(np.isnan(r).sum()),
                                 int(inside.shape[0] > cap), cap], dtype=torch.int64)
        return cand, counters

    # fused small selects: numpy versions of k_small_select / k_cand_hist / k_cand_scan / k_cand_finish
    NBINS, FINCAP = 4096, 4096

    def small_select(self, keys, seg_counts, nseg, seg_cap, ra, rb, npass, use_scan=False, read=False):
        k = keys.numpy().reshape(nseg, seg_cap)
        cnts = [seg_cap] * nseg if seg_counts is None else [int(c) for c in seg_counts.tolist()]
        valid = True
        if use_scan:
            if self.fs["is_nan"]:
                out = (float("nan"),) * 3 + (True,)
                self.last = torch.tensor([out[0], out[1], out[2], 1.], dtype=torch.float64)
                return out if read else self.last
            valid = self.fs["valid"] and all(c <= seg_cap for c in cnts)
            ra, rb = self.fs["r0"] - self.fs["base_a"], self.fs["r1"] - self.fs["base_a"]
        v = np.concatenate([k[i, :min(cnts[i], seg_cap)] for i in range(nseg)]) if nseg else np.zeros(0)
        v = np.sort(v[~np.isnan(v)])
        if valid and rb >= v.size:
            valid = False
        if not valid:
            a = b = 0.
        else:
            a, b = float(v[ra]), float(v[rb])
            if npass == 3:                                  # bracket: round outwards over 25 bits
                ua = np.array([a]).view(np.uint64)[0] >> np.uint64(25) << np.uint64(25)
                ub = np.array([b]).view(np.uint64)[0] | np.uint64((1 << 25) - 1)
                a = float(np.array([ua], dtype=np.uint64).view(np.float64)[0])
                b = float(np.array([ub], dtype=np.uint64).view(np.float64)[0])
        self.last = torch.tensor([(a + b) / 2. * 2., a, b, 1. if valid else 0.], dtype=torch.float64)
        if not read:
            return self.last
        return (a + b) / 2. * 2., a, b, valid

    @staticmethod
    def _bins(r, lo, hi):
        w = hi - lo
        scale = NumpySelect.NBINS / w if w > 0 else 0.
        t = (r - lo) * scale
        return np.clip(np.where(t > 0, np.minimum(t, NumpySelect.NBINS - 1), 0), 0, NumpySelect.NBINS - 1).astype(np.int64)

    def cand_hist(self, cand, count, lohi):
        n = min(int(count.item()), cand.shape[0])
        b = self._bins(cand.numpy()[:n], float(lohi[1]), float(lohi[2]))
        return torch.from_numpy(np.bincount(b, minlength=self.NBINS).astype(np.int32))

    def cand_scan(self, fhist, counters, k0, k1):
        below, ncand, nan, over, cap_total = [int(v) for v in counters.tolist()]
        ok = nan == 0 and over == 0 and ncand <= cap_total and k0 >= below and k1 < below + ncand
        r0, r1 = (k0 - below, k1 - below) if ok else (0, 0)
        c = np.cumsum(fhist.numpy().astype(np.int64))
        ba = int(np.searchsorted(c, r0, side="right"))
        bb = int(np.searchsorted(c, r1, side="right"))
        self.fs = dict(valid=bool(ok or nan), is_nan=bool(nan), r0=r0, r1=r1, bin_a=ba, bin_b=bb,
                       base_a=int(c[ba - 1]) if ba > 0 else 0)

    def cand_gather(self, cand, count, lohi):
        fin = torch.zeros(self.FINCAP, dtype=torch.float64)
        if not self.fs["valid"] or self.fs["is_nan"]:
            return fin, torch.zeros(1, dtype=torch.int32)
        n = min(int(count.item()), cand.shape[0])
        r = cand.numpy()[:n]
        b = self._bins(r, float(lohi[1]), float(lohi[2]))
        pick = r[(b >= self.fs["bin_a"]) & (b <= self.fs["bin_b"])]
        fin[:min(pick.size, self.FINCAP)] = torch.from_numpy(pick[:self.FINCAP].copy())
        return fin, torch.tensor([pick.size], dtype=torch.int32)

    def begin_bracket(self, k0, k1, counters):
        below, ncand, nan, over, cap_total = [int(v) for v in counters.tolist()]
        ok = over == 0 and ncand <= cap_total and k0 >= below and k1 < below + ncand
        self.prefix = [0, 0]
        self.rank = [k0 - below, k1 - below] if ok else [0, 0]
        self.nprefix = 1
        self.valid = ok
        self.hist.zero_()
        self.nan.zero_()
        self.nan += nan


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, with_nan, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from pyxfocus_b200 import dist
    rng = np.random.default_rng(seed)
    x = rng.normal(3., 1e-3, n)
    y = rng.standard_cauchy(n) * 1e-3 - 1.
    x[: n // 20] = x[0]
    y[: n // 20] = y[0]                                  # exact ties across shards
    if with_nan:
        x[n // 2] = np.nan
    lo, hi = dist.shard_range(n, rank, world)
    xs, ys = x[lo:hi], y[lo:hi]
    # centroid: all-reduce of (count, sum x, sum y) exactly as dist.hpd does with the CUDA sums
    s = torch.tensor([float(hi - lo), xs.sum(), ys.sum()], dtype=torch.float64)
    dist.all_reduce_sum(s)
    cx, cy = float(s[1] / s[0]), float(s[2] / s[0])
    sel = NumpySelect(xs, ys, cx, cy)
    m = torch.tensor([float(hi - lo)], dtype=torch.float64)
    td.all_reduce(m, op=td.ReduceOp.MIN)
    res = dist.bracket_median_pair(sel, n, int(m.item()))
    used_bracket = res is not None and res[3]
    if not used_bracket:
        res = dist.select_median_pair(sel, n)
    q.put((rank, res[:3], cx, cy, bool(used_bracket), sel.collect_calls))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("world,n,with_nan", [(2, 100_001, False), (3, 4_100, False), (2, 60_000, True)])
def test_sharded_exact_median_over_gloo(world, n, with_nan):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 123, with_nan, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the whole bundle
    rng = np.random.default_rng(123)
    x = rng.normal(3., 1e-3, n)
    y = rng.standard_cauchy(n) * 1e-3 - 1.
    x[: n // 20] = x[0]
    y[: n // 20] = y[0]
    if with_nan:
        x[n // 2] = np.nan
    res0 = out[0][1]
    for rank, res, cx, cy, used, calls in out:
        assert res == res0 or (np.isnan(res[0]) and np.isnan(res0[0])), "ranks disagree"
        if n >= 50_000 and not with_nan:
            assert used and calls == 1, "bracket path (one pass over the shard) not taken"
    cx, cy = out[0][2], out[0][3]
    r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
    if with_nan:
        assert np.isnan(res0[0])
    else:
        assert res0[0] == 2. * np.median(r)
        srt = np.sort(r)
        assert res0[1] == srt[(n - 1) // 2] and res0[2] == srt[n // 2]


# ---------------------------------------------------------------- weighted HPD
class NumpyWeighted:
    """CPU stand-in for dist.CudaWeighted: sorted radius keys + prefix sums of the weights."""

    def __init__(self, x, y, w, cx, cy):
        r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
        idx = np.argsort(r, kind="stable")
        self.keys = torch.from_numpy(np.ascontiguousarray(r[idx]).view(np.int64).copy())
        self.cum = torch.from_numpy(np.cumsum(w[idx]))
        self.device = torch.device("cpu")


def _weighted_bundle(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(3., 1e-3, n)
    y = rng.standard_cauchy(n) * 1e-3 - 1.
    w = rng.uniform(.2, 3., n)
    x[: n // 50] = x[0]
    y[: n // 50] = y[0]                                  # ties across shards
    return x, y, w


def _weighted_worker(rank, world, port, n, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from pyxfocus_b200 import dist
    x, y, w = _weighted_bundle(n, seed)
    lo, hi = dist.shard_range(n, rank, world)
    if rank == world - 1 and world == 3:
        lo = hi                                          # an EMPTY shard must not break the merge
    xs, ys, ws = x[lo:hi], y[lo:hi], w[lo:hi]
    s = torch.tensor([ws.sum(), (ws * xs).sum(), (ws * ys).sum()], dtype=torch.float64)
    dist.all_reduce_sum(s)
    cx, cy = float(s[1] / s[0]), float(s[2] / s[0])
    loc = NumpyWeighted(xs, ys, ws, cx, cy)
    wl = loc.cum[-1].clone() if loc.cum.shape[0] else torch.zeros((), dtype=torch.float64)
    W = dist.all_reduce_sum(wl)
    r75 = float(dist.weighted_quantile_radius(loc, .75, W))
    r25 = float(dist.weighted_quantile_radius(loc, .25, W))
    q.put((rank, r75, r25, cx, cy, lo, hi))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 40_001), (3, 9_000)])
def test_sharded_weighted_hpd_over_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_weighted_worker, args=(r, world, port, n, 321, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, w = _weighted_bundle(n, 321)
    covered = np.zeros(n, dtype=bool)
    for rank, r75, r25, cx, cy, lo, hi in out:
        covered[lo:hi] = True
        assert (r75, r25) == (out[0][1], out[0][2]), "ranks disagree"
    x, y, w = x[covered], y[covered], w[covered]        # (world 3 drops the last shard on purpose)
    cx, cy = out[0][3], out[0][4]
    r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
    ind = np.argsort(r)
    cdf = np.cumsum(w[ind])
    cdf = cdf / cdf.max()
    want75 = r[ind][np.argmin(np.abs(cdf - .75))]
    want25 = r[ind][np.argmin(np.abs(cdf - .25))]
    assert out[0][1] == want75 and out[0][2] == want25   # analyses.py:88-97, bit for bit


# ---------------------------------------------------------------- bracketed weighted HPD (pxf_wquant.cu logic)
class NumpyWeightedBracket:
    """CPU stand-in for dist.CudaWeightedBracket: same interface, numpy restatement of the pxf_wq_* kernels
    (k_wq_sample, k_wq_brackets, k_wq_collect) so that the sharded driver runs over gloo."""
    Z = 6.0

    def __init__(self, x, y, w, cx, cy, nsamp_total):
        self.r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
        self.w = w
        self.num = x.shape[0]
        self.device = torch.device("cpu")
        self.nsamp_total = nsamp_total

    def params(self, total):
        return 0, self.nsamp_total

    def sample(self, nsamp):
        out = np.zeros((2, nsamp))
        out[0] = np.inf
        take = min(nsamp, self.num)
        if take:
            idx = (np.arange(take, dtype=np.uint64) * np.uint64(self.num)) // np.uint64(take)
            out[0, :take] = self.r[idx.astype(np.int64)]
            out[1, :take] = self.w[idx.astype(np.int64)] * (self.num / take)
        return torch.from_numpy(out)

    def set_brackets(self, gathered):
        g = gathered.numpy()
        r, w = g[:, 0, :].reshape(-1), g[:, 1, :].reshape(-1)
        idx = np.argsort(r, kind="stable")
        rs, cum = r[idx], np.cumsum(w[idx])
        n, W = rs.shape[0], cum[-1]
        wi = np.diff(cum, prepend=0.)
        design = max(n * float((wi * wi).sum()) / (W * W), 1.)
        self.lohi = []
        for q in (.25, .75):
            delta = self.Z * np.sqrt(q * (1 - q) * design / n) + 2. / n
            lo, hi = 0., np.inf
            if q - delta > 0:
                p = int(np.searchsorted(cum, (q - delta) * W, side="left"))
                lo = rs[p - 1] if p > 0 else 0.
            if q + delta < 1:
                p = int(np.searchsorted(cum, (q + delta) * W, side="left"))
                hi = rs[p + 1] if p + 1 < n else np.inf
            self.lohi.append((lo, hi))

    def collect(self, cap):
        out = np.zeros(5)
        self.win = []
        for b, (lo, hi) in enumerate(self.lohi):
            out[b] = self.w[self.r < lo].sum()
            inside = (self.r >= lo) & (self.r <= hi)
            out[2 + b] = max(int(inside.sum()) - cap, 0)
            self.win.append((self.r[inside], self.w[inside]))
        out[4] = int(np.isnan(self.r).sum() + (~(self.w >= 0)).sum())
        return torch.from_numpy(out)

    def windows(self):
        res = []
        for r, w in self.win:
            idx = np.argsort(r, kind="stable")
            res.append((torch.from_numpy(np.ascontiguousarray(r[idx]).view(np.int64).copy()),
                        torch.from_numpy(np.cumsum(w[idx]))))
        return res


def _bracket_worker(rank, world, port, n, seed, nsamp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from pyxfocus_b200 import dist
    x, y, w = _weighted_bundle(n, seed)
    lo, hi = dist.shard_range(n, rank, world)
    if rank == world - 1 and world == 3:
        lo = hi                                          # an EMPTY shard
    xs, ys, ws = x[lo:hi], y[lo:hi], w[lo:hi]
    s = torch.tensor([ws.sum(), (ws * xs).sum(), (ws * ys).sum(), float(hi - lo)], dtype=torch.float64)
    dist.all_reduce_sum(s)
    cx, cy = float(s[1] / s[0]), float(s[2] / s[0])
    loc = NumpyWeightedBracket(xs, ys, ws, cx, cy, nsamp)
    res, ok = dist.hpd_weighted_bracketed(loc, int(s[3]), s[0])
    q.put((rank, float(res), ok, cx, cy, lo, hi))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("world,n,nsamp", [(2, 200_001, 16384), (3, 90_000, 8192)])
def test_sharded_bracketed_weighted_hpd_over_gloo(world, n, nsamp):
    """dist.hpd_weighted_bracketed: gathered sample -> brackets -> local collect -> 64-ary merge of the sorted
    candidate windows, over gloo at world 2 and 3 (one empty shard): same radii as numpy's full
    argsort -> cumsum -> argmin on the whole bundle."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bracket_worker, args=(r, world, port, n, 654, nsamp, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, w = _weighted_bundle(n, 654)
    covered = np.zeros(n, dtype=bool)
    for rank, res, ok, cx, cy, lo, hi in out:
        covered[lo:hi] = True
        assert ok, "bracket reported a miss"
        assert res == out[0][1], "ranks disagree"
    x, y, w = x[covered], y[covered], w[covered]
    cx, cy = out[0][3], out[0][4]
    r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
    ind = np.argsort(r)
    cdf = np.cumsum(w[ind])
    cdf = cdf / cdf.max()
    want = r[ind][np.argmin(np.abs(cdf - .75))] - r[ind][np.argmin(np.abs(cdf - .25))]
    assert out[0][1] == pytest.approx(want, rel=1e-12)


def test_kary_merge_matches_bisection_semantics():
    """weighted_quantile_radii on one rank == numpy argmin semantics for random runs, including the
    predecessor-is-closer case, an offset, and validity flags."""
    from pyxfocus_b200 import dist
    rng = np.random.default_rng(77)
    for trial in range(20):
        n = int(rng.integers(1, 400))
        r = np.sort(rng.random(n))
        w = rng.random(n) + .01
        cum = np.cumsum(w)
        off = float(rng.random() * 2.) if trial % 2 else 0.
        W = off + cum[-1] + (float(rng.random()) if trial % 3 == 0 else 0.)
        qs = [.25, .75]
        runs = [(torch.from_numpy(r.view(np.int64).copy()), torch.from_numpy(cum))] * 2
        got, valid = dist.weighted_quantile_radii(runs, qs, torch.tensor(W), offsets=torch.tensor([off, off]))
        cdf = (off + cum) / W
        for i, qq in enumerate(qs):
            reach = np.nonzero(cdf >= qq)[0]
            if reach.size == 0:
                assert not bool(valid[i])
                continue
            j = int(reach[0])
            if j == 0:
                assert bool(valid[i]) == (off == 0.)
                want = r[0]
            else:
                assert bool(valid[i])
                want = r[j - 1] if abs(cdf[j - 1] - qq) <= abs(cdf[j] - qq) else r[j]
            if bool(valid[i]):
                assert float(got[i]) == want


# ---------------------------------------------------------------- sharded rhocdf: sample sort + all-to-all
class NumpySortedRun:
    """CPU stand-in for dist.CudaSortedRun."""

    def __init__(self, rays, weights, cx, cy):
        x, y = rays[1].numpy(), rays[2].numpy()
        r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
        idx = np.argsort(r, kind="stable")
        self.r = torch.from_numpy(np.ascontiguousarray(r[idx]))
        w = np.ones_like(r) if weights is None else np.asarray(weights, dtype=np.float64)
        self.w = torch.from_numpy(np.ascontiguousarray(w[idx]))
        self.device = torch.device("cpu")

    @staticmethod
    def merge(r, w):
        idx = np.argsort(r.numpy(), kind="stable")
        return torch.from_numpy(r.numpy()[idx].copy()), torch.from_numpy(w.numpy()[idx].copy())

    @staticmethod
    def prefix(w):
        return torch.from_numpy(np.cumsum(w.numpy()))


def _rhocdf_worker(rank, world, port, n, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from pyxfocus_b200 import dist
    x, y, w = _weighted_bundle(n, seed)
    lo, hi = dist.shard_range(n, rank, world)
    if rank == 1 and world == 3:
        hi = lo                                          # an EMPTY shard in the middle
    rays = [None, torch.from_numpy(x[lo:hi].copy()), torch.from_numpy(y[lo:hi].copy())] + [None] * 7
    # cent=False: radii about the origin (the centroid path needs the CUDA sums kernel)
    r, cdf, first = dist.rhocdf(rays, weights=w[lo:hi], cent=False, local_cls=NumpySortedRun)
    q.put((rank, r.numpy(), cdf.numpy(), first, lo, hi))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 30_001), (3, 12_000)])
def test_sharded_rhocdf_sample_sort_over_gloo(world, n):
    """dist.rhocdf: local sort -> agreed splitters -> one all-to-all of (radius, weight) pairs -> merge -> prefix sums
    with all-gathered offsets.  The slices, concatenated in rank order, are numpy's argsort -> cumsum -> /max of the
    whole bundle (analyses.py:73-86); ties across shards and an empty shard included."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rhocdf_worker, args=(r, world, port, n, 99, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, w = _weighted_bundle(n, 99)
    covered = np.zeros(n, dtype=bool)
    for _, _, _, _, lo, hi in out:
        covered[lo:hi] = True
    x, y, w = x[covered], y[covered], w[covered]
    r = np.sqrt(x ** 2 + y ** 2)
    ind = np.argsort(r, kind="stable")
    cdf = np.cumsum(w[ind])
    cdf = cdf / cdf.max()
    got_r = np.concatenate([t[1] for t in out])
    got_c = np.concatenate([t[2] for t in out])
    assert np.array_equal(got_r, r[ind])
    assert np.abs(got_c - cdf).max() <= 1e-13
    firsts = [t[3] for t in out]
    assert firsts == list(np.cumsum([0] + [t[1].shape[0] for t in out[:-1]]))
    assert sum(t[1].shape[0] for t in out) == int(covered.sum())
