"""Golden fixture for the wavefront-reconstruction row: tests/golden/southwell.npz.

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden_southwell.py

Executes the reference's *unmodified* ``southwell.py`` (padArrays, southwell; its ``example()`` input) with the
C oracle standing in for the f2py module ``reconstruct`` (no Fortran compiler here), and the oracle's
``southwellbin`` on a traced bundle binned into lenslets, followed by the reference's own southwell()."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import chains, f2py as of, pyref, refload  # noqa: E402


def load_reference_southwell():
    rec = types.ModuleType("reconstruct")
    rec.__dict__.update(vars(of.reconstruct))
    sys.modules["reconstruct"] = rec
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    spec = importlib.util.spec_from_file_location("ref_southwell", os.path.join(refload.REFERENCE_ROOT, "southwell.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    sw = load_reference_southwell()
    out = {}
    # ---- southwell.example() input at 64 x 64 (southwell.py:53-63)
    n = 64
    xg, yg = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-1, 1, n))
    img = np.polynomial.legendre.legval2d(xg, yg, [[0, 1, 0], [0, .5, 0], [1, 0, 0]])
    gx, gy = np.gradient(img)
    rad = np.sqrt(xg ** 2 + yg ** 2)
    gx[rad > 1] = np.nan
    gy[rad > 1] = np.nan
    out["ex_gx"], out["ex_gy"] = gx.copy(), gy.copy()
    out["ex_phase"] = sw.southwell(gx, gy, 1e-10, 1., maxiter=10000)
    out["ex_sweeps"] = of.reconstruct.reconstruct.sweeps
    # ---- an irregular aperture with isolated lenslets (exercises the goodpix == 0 invalidation, reconstruct.f95:88-94)
    rng = np.random.default_rng(7)
    gx = rng.normal(0., 1e-3, (40, 50))
    gy = rng.normal(0., 1e-3, (40, 50))
    hole = rng.random((40, 50)) < .35
    gx[hole] = np.nan
    gy[hole] = np.nan
    out["ir_gx"], out["ir_gy"] = gx.copy(), gy.copy()
    out["ir_phase"] = sw.southwell(gx, gy, 1e-10, 1., maxiter=300)
    out["ir_sweeps"] = of.reconstruct.reconstruct.sweeps
    # ---- lenslet binning of a traced bundle (reconstruct.f95:136-187) + reconstruction
    rays = chains.wolter1_source(20000, seed=3, dphi=.4)
    pyref.transform(rays, 220.3, 0, 0, 0, 0, 0)
    rays[4] = rays[4] + 1e-4 * rays[1]                      # a focusing wavefront: slopes proportional to position
    rays[5] = rays[5] - 2e-4 * rays[2]
    out["bin_x"], out["bin_y"], out["bin_l"], out["bin_m"] = rays[1], rays[2], rays[4], rays[5]
    for tag, (xd, yd, bs) in (("even", (12, 30, 3.2)), ("odd", (11, 31, 3.2))):
        xa, ya, ph = of.reconstruct.southwellbin(rays[1], rays[2], rays[4], rays[5], bs, xd, yd)
        out["bin_%s_xang" % tag], out["bin_%s_yang" % tag], out["bin_%s_phase" % tag] = xa.copy(), ya.copy(), ph.copy()
        out["bin_%s_dims" % tag] = np.array([xd, yd, bs])
        pc = of.reconstruct.reconstruct(xa, ya, 1e-12, bs, ph, 2000)
        out["bin_%s_phasec" % tag] = pc
        out["bin_%s_sweeps" % tag] = of.reconstruct.reconstruct.sweeps
    np.savez_compressed(os.path.join(HERE, "southwell.npz"), **out)
    print("southwell.npz:", {k: np.shape(v) for k, v in out.items()})
    print("sweeps:", out["ex_sweeps"], out["ir_sweeps"], out["bin_even_sweeps"], out["bin_odd_sweeps"])


if __name__ == "__main__":
    main()
