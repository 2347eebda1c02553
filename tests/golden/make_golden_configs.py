"""Golden fixtures for the WHOLE chains of BASELINE configs 2-5 (config 1: wolter1.npz, make_golden.py).

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden_configs.py

The scripts in ``pyxfocus_b200/examples.py`` are executed on the reference's own, unmodified Python layer
(``oracle.refload``: sources / transformations / surfaces / analyses imported from /root/reference, with the C
oracle in the four f2py slots because no Fortran compiler exists here).  ``findimageplane`` has no definition in
the reference; the literal scan in examples.py stands in for it (parity unpinned).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refload  # noqa: E402
from pyxfocus_b200 import examples as ex  # noqa: E402  (imports libpxf.so; no GPU work happens here)

SIZES = dict(c2_n=3000, c2_arcmin=(0., 5., 10., 20., 30.), c3_n=4000, c4_n=20, c4_M=72, c5_n=100, c5_shells=40)


def stack(rays, rows):
    return np.stack([np.asarray(rays[k], dtype=np.float64) for k in rows])


def main():
    fortran = "--fortran-source" in sys.argv      # the reference's Fortran TEXT in the f2py slots: compare, write nothing
    if fortran:
        from oracle import f95mods
        api = ex.make_api(refload.load(f2py_modules=f95mods.modules()), ex.NumpyXP, "reference over its own Fortran")
    else:
        api = ex.make_api(refload.load(), ex.NumpyXP, "reference")
    g = {}
    ap = ex.ws_aperture(api)
    g["c2_aperture"] = np.array(ap)
    for a in SIZES["c2_arcmin"]:
        r = ex.config2_point(api, SIZES["c2_n"], a / 60. * np.pi / 180., ap)
        tag = "c2_%02d_" % int(a)
        g[tag + "xy"] = stack(r["rays"], (1, 2))
        g[tag + "scalars"] = np.array([r["f"], r["d2"], r["d3"], r["hpd"], r["rms"], r["hpd_scan"], r["rms_scan"]])
    r = ex.config3(api, SIZES["c3_n"])
    g["c3_rows"] = stack(r["rays"], range(10))
    g["c3_idx"] = np.asarray(r["idx"], dtype=np.int64)
    g["c3_scalars"] = np.array([r["hpd"], r["rms"]])
    # (beyond |order * wave| = 7.2 nm every ray of this geometry is evanescent: higher orders take shorter wavelengths)
    for order, wave in ((-1, 4.8), (-3, 2.4), (-8, .6), (-1, "uniform")):
        r = ex.config4(api, SIZES["c4_n"], SIZES["c4_M"], order=order, wave=wave)
        tag = "c4_o%d_%s_" % (-order, "w" if isinstance(wave, str) else "s")
        g[tag + "rows"] = stack(r["rays"], range(1, 10))
        g[tag + "scalars"] = np.array([r["kept"], r["dz"], r["gratings"], r["cx"], r["cy"], r["rmsY"], r["hpdY"]])
    r = ex.config5(api, SIZES["c5_n"], SIZES["c5_shells"], offaxis=1. / 60. * np.pi / 180.)
    g["c5_rows"] = stack(r["rays"], range(1, 10))
    g["c5_weights"] = np.asarray(r["weights"])
    g["c5_scalars"] = np.array([r["kept"], r["hpd"], r["rms"], r["cx"], r["cy"], r["area"]])
    if fortran:
        old = np.load(os.path.join(HERE, "configs.npz"))
        bad = [k for k in g if k not in old.files or not np.array_equal(np.asarray(g[k]), old[k], equal_nan=True)]
        print("%-18s %3d arrays: %s" % ("configs.npz", len(g), "identical" if not bad else "DIFFER: " + ", ".join(bad)))
        return
    np.savez_compressed(os.path.join(HERE, "configs.npz"), **g)
    print("wrote configs.npz:", sum(v.nbytes for v in g.values()), "bytes raw")


if __name__ == "__main__":
    main()
