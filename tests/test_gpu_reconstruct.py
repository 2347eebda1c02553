"""GPU parity for the wavefront-reconstruction row: pyxfocus_b200.reconstruct /
.southwell against the golden fixture (reference southwell.py + oracle) and against the oracle directly.
The Gauss-Seidel pipeline must reproduce the sequential loop bit for bit, sweep count included."""
import numpy as np
import pytest

from oracle import f2py as of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pxf():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pyxfocus_b200
    return pyxfocus_b200


def _pad100(a):
    t = np.zeros((a.shape[0] + 2, a.shape[1] + 2), order="F") + 100.
    t[1:-1, 1:-1] = a
    return t


def test_southwell_matches_the_reference_script(pxf, golden):
    from pyxfocus_b200 import reconstruct, southwell
    g = golden("southwell")
    before = pxf.launch_count()
    for tag, maxiter in (("ex", 10000), ("ir", 300)):
        gx, gy = g[tag + "_gx"].copy(), g[tag + "_gy"].copy()
        got = southwell.southwell(gx, gy, 1e-10, 1., maxiter=maxiter)
        assert reconstruct.reconstruct.sweeps == int(g[tag + "_sweeps"])
        assert np.array_equal(got, g[tag + "_phase"], equal_nan=True), "%s: phase differs" % tag
        assert (gx[np.isnan(g[tag + "_gx"])] == 100.).all()          # inputs modified in place like the reference
    assert pxf.launch_count() > before


@pytest.mark.parametrize("shape,fill,maxiter", [((3, 3), .0, 50), ((4, 7), .0, 500), ((65, 33), .2, 400), ((130, 171), .1, 3000),
                                                ((20, 20), 1., 5)])
def test_reconstruct_bit_identical_to_the_oracle(pxf, shape, fill, maxiter):
    """Random slopes with random holes (isolated lenslets get invalidated in-sweep), several shapes incl. the
    smallest interior, an all-invalid array (rms = 0/0: runs to maxiter) and a run that hits the cap."""
    from pyxfocus_b200 import reconstruct
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    gx = rng.normal(0., 1e-3, shape)
    gy = rng.normal(0., 1e-3, shape)
    hole = rng.random(shape) < fill
    gx[hole] = 100.
    gy[hole] = 100.
    ph = np.zeros(shape, order="F")
    ph[hole] = 100.
    a = [_pad100(v) for v in (gx, gy, ph)]
    b = [v.copy(order="F") for v in a]
    want = of.reconstruct.reconstruct(a[0], a[1], 1e-9, .7, a[2], maxiter)
    got = reconstruct.reconstruct(b[0], b[1], 1e-9, .7, b[2], maxiter)
    assert reconstruct.reconstruct.sweeps == of.reconstruct.reconstruct.sweeps
    assert np.array_equal(got, want)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)                                    # xang, yang, phase updated in place alike
    with pytest.raises(ValueError):
        reconstruct.reconstruct(np.ascontiguousarray(b[0]), b[1], 1e-9, .7, b[2], 3)   # C order: f2py refuses it


def test_southwellbin_and_reconstruct_on_a_traced_bundle(pxf, golden):
    import torch
    from pyxfocus_b200 import reconstruct
    g = golden("southwell")
    rows = [torch.from_numpy(g[k]).cuda() for k in ("bin_x", "bin_y", "bin_l", "bin_m")]
    for tag in ("even", "odd"):
        xd, yd, bs = g["bin_%s_dims" % tag]
        xa, ya, ph = reconstruct.southwellbin(*rows, bs, int(xd), int(yd))
        wx, wy, wp = g["bin_%s_xang" % tag], g["bin_%s_yang" % tag], g["bin_%s_phase" % tag]
        assert np.array_equal(ph, wp)                                   # same empty lenslets
        assert np.array_equal(xa == 100., wx == 100.)
        ok = wx != 100.
        assert np.abs(xa[ok] - wx[ok]).max() <= 1e-12 and np.abs(ya[ok] - wy[ok]).max() <= 1e-12
        # reconstruction from the oracle's slopes (identical inputs): same bits, same sweeps
        a = [np.asfortranarray(v.copy()) for v in (wx, wy, wp)]
        got = reconstruct.reconstruct(a[0], a[1], 1e-12, bs, a[2], 2000)
        assert reconstruct.reconstruct.sweeps == int(g["bin_%s_sweeps" % tag])
        assert np.array_equal(got, g["bin_%s_phasec" % tag])
    # numpy rows are accepted as well, and an empty bundle flags every lenslet
    xa, ya, ph = reconstruct.southwellbin(g["bin_x"][:0], g["bin_y"][:0], g["bin_l"][:0], g["bin_m"][:0], 1., 4, 5)
    assert (xa == 100.).all() and (ph == 100.).all()
