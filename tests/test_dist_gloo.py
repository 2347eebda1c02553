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
    """CPU stand-in for dist.CudaSelect (same interface, same semantics)."""

    def __init__(self, r):
        self.keys = np.ascontiguousarray(r, dtype=np.float64).view(np.uint64)
        self.isnan = np.isnan(r)
        self.hist = torch.zeros(2 * 8192, dtype=torch.int64)
        self.nan = torch.zeros(1, dtype=torch.int64)

    def schedule(self):
        return SCHEDULE

    def begin(self, k0, k1):
        self.prefix = [0, 0]
        self.rank = [int(k0), int(k1)]
        self.nprefix = 1
        self.hist.zero_()
        self.nan.zero_()

    def histogram(self, shift, bits):
        nb = 1 << bits
        k = self.keys[~self.isnan]
        self.nan += int(self.isnan.sum())
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
            newp.append(((p << bits) | b) & (2 ** 64 - 1))
            self.rank[j] -= base
        self.prefix = newp
        self.nprefix = 1 if newp[0] == newp[1] else 2
        self.hist.zero_()

    def nan_count(self):
        return self.nan

    def finish(self, total):
        a = np.array([self.prefix[0]], dtype=np.uint64).view(np.float64)[0]
        b = np.array([self.prefix[1]], dtype=np.uint64).view(np.float64)[0]
        med = (a + b) / 2.
        if total == 0 or int(self.nan.item()) > 0:
            med = float("nan")
        return 2. * med, float(a), float(b)


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
    x[: n // 4] = x[0]
    y[: n // 4] = y[0]                                   # exact ties across shards
    if with_nan:
        x[n // 2] = np.nan
    lo, hi = dist.shard_range(n, rank, world)
    xs, ys = x[lo:hi], y[lo:hi]
    # centroid: all-reduce of (count, sum x, sum y) exactly as dist.hpd does with the CUDA sums
    s = torch.tensor([float(hi - lo), xs.sum(), ys.sum()], dtype=torch.float64)
    dist.all_reduce_sum(s)
    cx, cy = float(s[1] / s[0]), float(s[2] / s[0])
    r = np.sqrt((xs - cx) ** 2 + (ys - cy) ** 2)
    res = dist.select_median_pair(NumpySelect(r), n)
    q.put((rank, res, cx, cy))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("world,n,with_nan", [(2, 100_001, False), (3, 4_100, False), (2, 1000, True)])
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
    x[: n // 4] = x[0]
    y[: n // 4] = y[0]
    if with_nan:
        x[n // 2] = np.nan
    res0 = out[0][1]
    for rank, res, cx, cy in out:
        assert res == res0 or (np.isnan(res[0]) and np.isnan(res0[0])), "ranks disagree"
    cx, cy = out[0][2], out[0][3]
    r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
    if with_nan:
        assert np.isnan(res0[0])
    else:
        assert res0[0] == 2. * np.median(r)
        srt = np.sort(r)
        assert res0[1] == srt[(n - 1) // 2] and res0[2] == srt[n // 2]
