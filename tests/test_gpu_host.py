"""GPU tests of the host-array entry point (pxf_host_trace_program): HOST rows in, HOST rows
mutated in place, bit-identical to the device-resident path and to the oracle."""
import numpy as np
import pytest

from util import assert_bit_equal, chains, copy, pyref, steps_to_program, to_dev, to_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pxf():
    import torch
    assert torch.cuda.is_available()
    import pyxfocus_b200
    return pyxfocus_b200


@pytest.mark.parametrize("n", [1, 7, 100_001, 5_000_001])
def test_host_trace_matches_oracle_and_device(pxf, n):
    """numpy (pageable) rows; n spans one partial chunk to several 2^21-ray chunks."""
    cpu = chains.wolter1_source(n, seed=50)
    host = copy(cpu)
    dev = to_dev(cpu)
    prog = steps_to_program(chains.wolter1_steps())
    out = pxf.host.trace(host, prog, write_back=True, hpd=True)
    prog.run(dev)
    if n <= 200_000:
        want_hpd = chains.wolter1_cpu(cpu)
        assert_bit_equal(host, cpu, rows=range(1, 10), what="host vs oracle")
        assert out["hpd"] == pytest.approx(want_hpd, rel=1e-9)
    assert_bit_equal(host, to_host(dev), rows=range(1, 10), what="host vs device")
    assert out["hpd"] == pxf.analyses.hpd(dev)
    assert out["alive_count"] == n


def test_host_trace_pinned_tensors_and_no_writeback(pxf):
    import torch
    n = 300_001
    cpu = chains.wolter1_source(n, seed=51)
    pinned = [None] + [torch.from_numpy(cpu[k].copy()).pin_memory() for k in range(1, 10)]
    before = [None] + [t.clone() for t in pinned[1:]]
    prog = steps_to_program(chains.wolter1_steps())
    out = pxf.host.trace(pinned, prog, write_back=False, hpd=True)         # opd row is None: never touched
    for k in range(1, 10):
        assert torch.equal(pinned[k], before[k]), "write_back=False must leave the host rows alone"
    ref = copy(cpu)
    assert out["hpd"] == pytest.approx(chains.wolter1_cpu(ref), rel=1e-9)
    out2 = pxf.host.trace(pinned, prog, write_back=True)
    assert "hpd" not in out2
    assert_bit_equal([np.zeros(n)] + [t.numpy() for t in pinned[1:]], ref, rows=range(1, 10), what="pinned rows")


def test_host_trace_vignette_program(pxf):
    n = 2_500_003
    cpu = chains.wolter1_source(n, seed=52, dphi=1.2)
    host = copy(cpu)
    prog = (pxf.Program().transform(0, 0, 8400., 0, 0, 0).wolterprimary(220., 8400., 1.).reflect()
            .vignette_box(3, 8426., 8526.).vignette_abs(2, 50.)
            .woltersecondary(220., 8400., 1.).reflect().flat())
    out = pxf.host.trace(host, prog, hpd=True, alive=True)
    dev = to_dev(cpu)
    alive = prog.run(dev)
    assert np.array_equal(out["alive"], alive.cpu().numpy())
    assert out["alive_count"] == int(alive.sum())
    assert_bit_equal(host, to_host(dev), rows=range(1, 10), what="vignette program rows")
    surv = pxf.transformations.vignette(dev, ind=alive.bool())
    assert out["hpd"] == pxf.analyses.hpd(surv)
    # oracle on a subsample: same survivors, same values
    sub = [r[:50_000].copy() for r in cpu]
    chains.run_steps_cpu(sub, chains.wolter1_steps()[:3])
    keep = (sub[3] > 8426.) & (sub[3] < 8526.) & (np.abs(sub[2]) < 50.)
    assert np.array_equal(out["alive"][:50_000].astype(bool), keep)


def test_host_trace_keep_xy_on_device(pxf):
    import torch
    n = 120_001
    cpu = chains.wolter1_source(n, seed=53)
    host = copy(cpu)
    keep = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
    pxf.host.trace(host, steps_to_program(chains.wolter1_steps()), keep_xy=keep)
    assert np.array_equal(keep[0].cpu().numpy(), host[1]) and np.array_equal(keep[1].cpu().numpy(), host[2])
    fake = [None, keep[0], keep[1]] + [keep[0]] * 7
    assert pxf.dist.hpd(fake) == pytest.approx(pyref.hpd(host), rel=1e-9)


def test_out_of_place_program_leaves_source_untouched(pxf):
    from pyxfocus_b200._call import bundle_alloc
    n = 50_001
    cpu = chains.wolter1_source(n, seed=54)
    src = to_dev(cpu)
    out = bundle_alloc(n, src[1].device, zero=True)
    steps_to_program(chains.wolter1_steps()).run(src, out=out)
    assert_bit_equal(to_host(src), cpu, what="source untouched")
    chains.run_steps_cpu(cpu, chains.wolter1_steps())
    assert_bit_equal(to_host(out), cpu, rows=range(1, 10), what="out-of-place result")


def test_sharded_hpd_single_rank_matches_analyses(pxf):
    """dist.hpd with world size 1 (no process group) == analyses.hpd, bracket and full paths."""
    for n in (10_001, 3_000_000):
        cpu = chains.wolter1_source(n, seed=55)
        dev = to_dev(cpu)
        steps_to_program(chains.wolter1_steps()).run(dev)
        assert pxf.dist.hpd(dev) == pxf.analyses.hpd(dev)
        assert pxf.dist.rmsCentroid(dev) == pytest.approx(pxf.analyses.rmsCentroid(dev), rel=1e-12)
        cx, cy = pxf.dist.centroid(dev)
        ax, ay = pxf.analyses.centroid(dev)
        assert (cx, cy) == (ax, ay)
        assert pxf.dist.analyticImagePlane(dev) == pytest.approx(pxf.analyses.analyticImagePlane(dev), rel=1e-12)


def test_fused_centroid_sums(pxf):
    """Program.run(..., sums=) returns the centroid sums of the final bundle from the trace
    kernel itself (specialised chain and generic interpreter) and hpd(..., sums=) uses them."""
    import torch
    for n in (1, 50_001, 3_000_001):
        cpu = chains.wolter1_source(n, seed=56)
        for steps in (chains.wolter1_steps(), chains.wolter1_steps()[:4]):          # chain kernel / interpreter
            dev = to_dev(cpu)
            ref = to_dev(cpu)
            sums = torch.zeros(16, dtype=torch.float64, device="cuda")
            prog = steps_to_program(steps)
            prog.run(dev, sums=sums)
            prog.run(ref)
            assert_bit_equal(to_host(dev), to_host(ref), what="sums variant changes no ray")
            s = sums.cpu().numpy()
            x, y = to_host(ref)[1], to_host(ref)[2]
            assert s[0] == n and s[3] == n
            assert s[1] == pytest.approx(x.sum(), rel=1e-11, abs=1e-9)
            assert s[2] == pytest.approx(y.sum(), rel=1e-11, abs=1e-9)
            h0 = pxf.analyses.hpd(ref)
            assert pxf.analyses.hpd(dev, sums=sums) == pytest.approx(h0, rel=1e-9)
    # with vignetting the sums cover the surviving rays only
    cpu = chains.wolter1_source(200_001, seed=57, dphi=1.2)
    dev = to_dev(cpu)
    sums = torch.zeros(16, dtype=torch.float64, device="cuda")
    prog = (pxf.Program().transform(0, 0, 8400., 0, 0, 0).wolterprimary(220., 8400., 1.).reflect()
            .vignette_abs(2, 50.).woltersecondary(220., 8400., 1.).reflect().flat())
    alive = prog.run(dev, sums=sums).bool()
    s = sums.cpu().numpy()
    assert s[0] == int(alive.sum())
    assert s[1] == pytest.approx(float(dev[1][alive].sum()), rel=1e-11, abs=1e-9)


def test_host_trace_constant_input_chunks_are_not_uploaded_but_still_exact(pxf, monkeypatch):
    """Input rows are scanned chunk by chunk; bitwise-constant chunks are filled on the device
    instead of uploaded.  Rows that are constant in some chunks only, constant -0 / NaN rows and
    the fully varying case must all give the bits of the plain path (PXF_HOST_NO_SCAN=1 is read
    once per process, so the comparison is against the device-resident program)."""
    n = 5_000_001                                   # three 2^21-ray chunks
    cpu = chains.wolter1_source(n, seed=54)         # z, l, m, n constant in every chunk
    # make l constant in the first chunk only, m = -0 everywhere, and one NaN ray in x
    cpu[4][(1 << 21):] += np.random.default_rng(1).normal(0., 1e-7, n - (1 << 21))
    cpu[5][:] = -0.
    cpu[6][:] = -np.sqrt(1. - cpu[4] ** 2)
    cpu[1][1234567] = np.nan
    host = copy(cpu)
    dev = to_dev(cpu)
    prog = steps_to_program(chains.wolter1_steps())
    pxf.host.trace(host, prog, write_back=True)
    prog.run(dev)
    want = to_host(dev)
    for k in range(1, 10):
        a, b = host[k], want[k]
        na, nb = np.isnan(a), np.isnan(b)
        assert np.array_equal(na, nb)
        assert np.array_equal(a.view(np.uint64)[~na], b.view(np.uint64)[~nb]), k


def test_host_trace_program_with_a_zernike_surface(pxf):
    """A program holding PXF_OP_ZERNSURF through the host-array entry point (its table is a device copy the
    Python layer uploads once): same bits as the device-resident program, opd row included."""
    from util import random_bundle
    n = 2_300_001                                            # two chunks
    ro, ao = chains.zernike_orders(7)
    coeff = chains.zernike_coeff(36, 5)
    cpu = random_bundle(n, 78)
    cpu[1] *= .4
    cpu[2] *= .4
    prog = (pxf.Program().transform(-1., 2., 100., 0, 0, .2).zernsurf(coeff, ro, ao, 62.5, 1.).reflect().flatopd(1.))
    dev = to_dev(cpu)
    prog.run(dev)
    host = copy(cpu)
    pxf.host.trace(host, prog, write_back=True)
    assert_bit_equal(host, to_host(dev), what="host vs device, Zernike program")
