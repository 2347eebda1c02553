"""Golden vectors from the reference's Fortran SOURCE TEXT, executed by oracle/f95run.py (a Fortran-subset translator with
gfortran's arithmetic rules; there is no Fortran compiler in the image).  Every subroutine of transformationsf.f95,
surfacesf.f95, woltsurf.f95 and zernsurf.f95 is run on seeded rays; inputs and outputs go to tests/golden/f95_source.npz.

    python tests/golden/make_f95_golden.py            # needs /root/reference; rewrites the fixture
    python tests/golden/make_f95_golden.py --check     # also prints the comparison with the C oracle

tests/test_f95_source.py holds oracle/pxf_oracle.c (and, on a GPU, libpxf) to these vectors bit for bit."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import chains, f95run, pyref  # noqa: E402
from oracle import f2py as of  # noqa: E402

REF = "/root/reference"
N = 40
ROWS = ["opd", "x", "y", "z", "l", "m", "n", "ux", "uy", "uz"]
HIDDEN = {"num", "arrsize", "arrsize1", "arrsize2", "cnum", "nc", "np"}      # f2py derives these from array lengths


def rays_above_primary(seed, dphi=.5):
    r = chains.wolter1_source(N, seed, dphi=dphi)
    pyref.transform(r, 0, 0, -8400. - 100., 0, 0, 0)        # start 100 mm above z0 so that surfaces at z ~ z0 are ahead
    pyref.transform(r, 0, 0, 100., 0, 0, 0)
    r[0] = np.random.default_rng(seed).normal(0., 1., N)
    return r


def tilted(seed, rx=1e-3, ry=-2e-3, dphi=.5):
    """Off-axis rays: every component of position and direction takes part in the Newton steps."""
    r = rays_above_primary(seed, dphi)
    pyref.transform(r, 0, 0, 0, rx, ry, 0)
    return r


def after_primary(seed):
    r = rays_above_primary(seed)
    of.woltsurf.wolterprimary(*r[1:], 220., 8400., 1.)
    of.transformationsf.reflect(*r[4:])
    return r


def generic(seed, spread=.02):
    """Rays from z ~ +50 heading down (-z) with a small angular spread, positions O(10)."""
    g = np.random.default_rng(seed)
    x, y = g.uniform(-8, 8, N), g.uniform(-8, 8, N)
    z = g.uniform(40, 60, N)
    l, m = g.normal(0, spread, N), g.normal(0, spread, N)
    n = -np.sqrt(1 - l ** 2 - m ** 2)
    u = g.normal(0, 1, (3, N))
    u /= np.sqrt((u ** 2).sum(0))
    return [g.normal(0, 1, N), x, y, z, l, m, n, u[0], u[1], u[2]]


def zern_tables(seed, nterms=21):
    ro = [r for r in range(8) for _ in range(r + 1)][:nterms]
    ao = [m for r in range(8) for m in range(-r, r + 1, 2)][:nterms]
    c = np.random.default_rng(seed).normal(0., 1e-3, nterms)
    c[:3] = 0.
    return c, np.array(ro, dtype=np.int32), np.array(ao, dtype=np.int32)


def ll_tables(seed):
    c = np.random.default_rng(seed).normal(0., 1e-5, 9)
    return c, np.repeat(np.arange(3), 3).astype(np.int32), np.tile(np.arange(3), 3).astype(np.int32)


def cases():
    """(file, fortran name, rays, {scalar/array arguments by their Fortran dummy name})"""
    alpha = pyref.woltparam(220., 8400.)[0]
    zc, zr, za = zern_tables(1)
    zc2, zr2, za2 = zern_tables(2, 10)
    lc, lax, laz = ll_tables(3)
    tilt = generic(5)
    out = [
        ("transformationsf", "reflect", generic(1), {}),
        ("transformationsf", "refract", after_primary(2), dict(n1=1., n2=1.5)),
        ("transformationsf", "transform", generic(3), dict(tx=1.5, ty=-2., tz=30., rx=.01, ry=-.02, rz=.3)),
        ("transformationsf", "transform", generic(3), dict(tx=0., ty=0., tz=-8400., rx=0., ry=0., rz=0.)),
        ("transformationsf", "itransform", generic(4), dict(tx=1.5, ty=-2., tz=30., rx=.01, ry=-.02, rz=.3)),
        ("transformationsf", "radgrat", generic(6), dict(wave=2.4e-6, dpermm=160. / 11832.911, order=-1.)),
        ("transformationsf", "radgratw", generic(6), dict(wave=np.random.default_rng(6).uniform(1e-6, 5e-6, N), dpermm=160. / 11832.911, order=-3.)),
        ("transformationsf", "grat", generic(7), dict(d=160e-6, order=np.repeat(-1., N), wave=np.random.default_rng(7).uniform(1e-6, 5e-6, N))),
        ("surfacesf", "flat", tilt, {}),
        ("surfacesf", "flatopd", generic(8), dict(nr=1.5)),
        ("surfacesf", "tracesphere", generic(9), dict(rad=500.)),
        ("surfacesf", "tracesphereopd", generic(9), dict(rad=500., nr=1.3)),
        ("surfacesf", "tracecyl", generic(10), dict(rad=300.)),
        ("surfacesf", "tracecylopd", generic(10), dict(rad=300., nr=1.3)),
        ("surfacesf", "cylconic", generic(11), dict(rad=300., k=-.5)),
        ("surfacesf", "conic", generic(12), dict(r=2e3, k=-1.)),
        ("surfacesf", "conic", generic(12), dict(r=2e3, k=-.3)),
        ("surfacesf", "conicopd", generic(13), dict(r=-2e3, k=-1.7, nr=1.2)),
        ("surfacesf", "paraxial", generic(14), dict(f=250.)),
        ("surfacesf", "paraxialy", generic(14), dict(f=-250.)),
        ("surfacesf", "torus", generic(15, .002), dict(rin=400., rout=900.)),
        ("surfacesf", "conicplus", generic(16), dict(r=2e3, k=-1., p=np.array([1e-9, -2e-12, 1e-15]))),
        ("surfacesf", "conicplusopd", generic(16), dict(r=2e3, k=-1., p=np.array([1e-9, -2e-12]), nr=1.1)),
        ("surfacesf", "legsurf", generic(17), dict(xwidth=10., ywidth=12., order=3, coeff=np.random.default_rng(17).normal(0, 1e-3, 6),
                                                    xo=np.array([0, 1, 2, 0, 1, 3], dtype=np.int32), yo=np.array([0, 0, 1, 2, 3, 1], dtype=np.int32))),
        ("woltsurf", "wolterprimary", rays_above_primary(20), dict(r0=220., z0=8400., psi=1.)),
        ("woltsurf", "wolterprimary", rays_above_primary(21), dict(r0=220., z0=8400., psi=2.5)),
        ("woltsurf", "wolterprimaryopd", rays_above_primary(22), dict(r0=220., z0=8400., psi=1., nr=1.)),
        ("woltsurf", "woltersecondary", after_primary(23), dict(r0=220., z0=8400., psi=1.)),
        ("woltsurf", "woltersine", rays_above_primary(24), dict(r0=220., z0=8400., amp=1e-4, freq=.05)),
        ("woltsurf", "wsprimary", rays_above_primary(25), dict(alpha=alpha, z0=8400., psi=1.)),
        ("woltsurf", "wsprimaryback", rays_above_primary(26), dict(alpha=alpha, z0=8400., psi=1., thick=.4)),
        ("woltsurf", "spocone", rays_above_primary(27), dict(r0=220., tg=.0065)),
        ("woltsurf", "wolterprimll", rays_above_primary(28, .3), dict(r0=220., z0=8400., zmax=8500., zmin=8400., dphi=.3, coeff=lc, axial=lax, az=laz)),
        ("woltsurf", "woltersecll", after_primary(29), dict(r0=220., z0=8400., psi=1., zmax=8400., zmin=8300., dphi=.5, coeff=lc, axial=lax, az=laz)),
        ("woltsurf", "ellipsoidwoltll", rays_above_primary(30), dict(r0=220., z0=8400., psi=1., s=1e5, zmax=8500., zmin=8400., dphi=.5, coeff=lc, axial=lax, az=laz)),
        ("zernsurf", "tracezern", generic(40, .002), dict(coeff=zc, rorder=zr, aorder=za, rad=12.)),
        ("zernsurf", "tracezernopd", generic(41, .002), dict(coeff=zc, rorder=zr, aorder=za, rad=12., nr=1.4)),
        ("zernsurf", "zernphase", generic(42, .002), dict(coeff=zc, rorder=zr, aorder=za, rad=12., wave=6e-4)),
        ("zernsurf", "tracezernrot", generic(43, .002), dict(coeff1=zc, rorder1=zr, aorder1=za, coeff2=zc2, rorder2=zr2, aorder2=za2, rad=12., rot=.3)),
    ]
    out += [
        ("woltsurf", "wolterprimary", tilted(50), dict(r0=220., z0=8400., psi=1.)),
        ("woltsurf", "wolterprimaryopd", tilted(51), dict(r0=220., z0=8400., psi=1., nr=1.5)),
        ("woltsurf", "woltersine", tilted(52), dict(r0=220., z0=8400., amp=1e-4, freq=.05)),
        ("woltsurf", "wsprimary", tilted(53), dict(alpha=alpha, z0=8400., psi=1.)),
        ("woltsurf", "wsprimary", tilted(54, 7e-3, 0.), dict(alpha=alpha, z0=8400., psi=1.)),      # 24 arcmin: beyond the graze angle
        ("woltsurf", "spocone", tilted(55), dict(r0=220., tg=.0065)),
        ("woltsurf", "wolterprimll", tilted(56, dphi=.3), dict(r0=220., z0=8400., zmax=8500., zmin=8400., dphi=.3, coeff=lc, axial=lax, az=laz)),
        ("woltsurf", "ellipsoidwoltll", tilted(57), dict(r0=220., z0=8400., psi=1., s=1e5, zmax=8500., zmin=8400., dphi=.5, coeff=lc, axial=lax, az=laz)),
    ]
    for seed, rx in ((58, 1e-3), (59, 7e-3)):
        r = tilted(seed, rx, -1e-3)
        of.woltsurf.wolterprimary(*r[1:], 220., 8400., 1.)
        of.transformationsf.reflect(*r[4:])
        out.append(("woltsurf", "woltersecondary", r, dict(r0=220., z0=8400., psi=1.)))
    # wssecondary / wssecondaryback need rays that left a W-S primary
    for name, extra, seed, rx in (("wssecondary", {}, 31, 0.), ("wssecondaryback", dict(thick=.4), 32, 0.),
                                  ("wssecondary", {}, 33, 1e-3), ("wssecondary", {}, 34, 7e-3)):
        r = tilted(seed, rx, 0.) if rx else rays_above_primary(seed)
        of.woltsurf.wsprimary(*r[1:], alpha, 8400., 1.)
        of.transformationsf.reflect(*r[4:])
        out.append(("woltsurf", name, r, dict(alpha=alpha, z0=8400., psi=1., **extra)))
    return out


def run_fortran(units, unit_args, rays, extra):
    """Call a translated subroutine with its full Fortran argument list; returns the ten rows after the call."""
    rows = {k: np.array(v, dtype=np.float64, copy=True) for k, v in zip(ROWS, rays)}
    args = []
    for a in unit_args:
        if a in extra:
            v = extra[a]
            args.append(np.array(v, copy=True) if isinstance(v, np.ndarray) else v)
        elif a in rows:
            args.append(rows[a])
        elif a == "num":
            args.append(N)
        elif a in ("arrsize", "arrsize1"):
            args.append(len(extra["coeff" if a == "arrsize" else "coeff1"]))
        elif a == "arrsize2":
            args.append(len(extra["coeff2"]))
        elif a == "cnum":
            args.append(len(extra["coeff"]))
        elif a == "nc":
            args.append(len(extra["coeff"]))
        elif a == "np":
            args.append(len(extra["p"]))
        else:
            raise KeyError("no value for dummy argument %s" % a)
    return args, rows


def run_oracle(module, name, unit_args, rays, extra):
    rows = {k: np.array(v, dtype=np.float64, copy=True) for k, v in zip(ROWS, rays)}
    args = []
    for a in unit_args:
        if a in HIDDEN:
            continue
        if a in extra:
            v = extra[a]
            args.append(np.array(v, copy=True) if isinstance(v, np.ndarray) else v)
        else:
            args.append(rows[a])
    getattr(getattr(of, module), name)(*args)
    return rows


def main():
    check = "--check" in sys.argv
    files = {m: f95run.load(os.path.join(REF, m + ".f95")) for m in ("transformationsf", "surfacesf", "woltsurf", "zernsurf")}
    sigs = {m: {u.name: u.args for u in f95run._parse_units(f95run._logical_lines(os.path.join(REF, m + ".f95"))).values()}
            for m in files}
    store, bad = {}, 0
    for k, (module, name, rays, extra) in enumerate(cases()):
        unit_args = sigs[module][name]
        args, rows = run_fortran(files[module], unit_args, rays, extra)
        files[module][name](*args)
        tag = "c%02d_%s_%s" % (k, module, name)
        store[tag + "__in"] = np.array(rays)
        store[tag + "__out"] = np.array([rows[r] for r in ROWS])
        for a, v in extra.items():
            store[tag + "__arg_" + a] = np.asarray(v)
        if check:
            orc = run_oracle(module, name, unit_args, rays, extra)
            diff = [r for r in ROWS if not np.array_equal(orc[r], rows[r], equal_nan=True)]
            worst = max((np.nanmax(np.abs(orc[r] - rows[r])) for r in diff), default=0.)
            changed = [r for r in ROWS if not np.array_equal(rows[r], np.asarray(rays[ROWS.index(r)]), equal_nan=True)]
            print("%-44s rows changed %-28s %s" % (tag, ",".join(changed), "== oracle" if not diff else "DIFFERS in %s (max %.3e)" % (",".join(diff), worst)))
            bad += bool(diff)
    # reconstruct.f95: the Southwell reconstructor and the lenslet binning
    rec = f95run.load(os.path.join(REF, "reconstruct.f95"))
    g = np.random.default_rng(70)
    for k2, (shape, holes, maxiter) in enumerate((((9, 7), 0, 200), ((14, 11), 12, 300))):
        yy, xx = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]))
        xang = np.asfortranarray(1e-3 * np.sin(xx / 3.) + 1e-5 * g.normal(size=shape))
        yang = np.asfortranarray(2e-3 * np.cos(yy / 4.) + 1e-5 * g.normal(size=shape))
        phase = np.zeros(shape, order="F")
        for _ in range(holes):
            i, j = g.integers(1, shape[0] - 1), g.integers(1, shape[1] - 1)
            xang[i, j] = yang[i, j] = phase[i, j] = 100.
        for a in (xang, yang, phase):
            a[0, :] = a[-1, :] = a[:, 0] = a[:, -1] = 100.
        fa, fb, fp = xang.copy(order="F"), yang.copy(order="F"), phase.copy(order="F")
        fc = np.zeros(shape, order="F")
        rec["reconstruct"](fa, fb, shape[0], shape[1], 1e-12, .5, fp, fc, maxiter)
        tag = "r%02d_reconstruct" % k2
        for nm, v in (("xang", xang), ("yang", yang), ("phase", phase), ("phasec_out", fc), ("phase_out", fp)):
            store[tag + "__" + nm] = v
        store[tag + "__maxiter"] = np.array(maxiter)
        if check:
            oa, ob, op = xang.copy(order="F"), yang.copy(order="F"), phase.copy(order="F")
            oc = of.reconstruct.reconstruct(oa, ob, 1e-12, .5, op, maxiter)
            same = np.array_equal(oc, fc) and np.array_equal(op, fp)
            print("%-44s %s" % (tag, "== oracle" if same else "DIFFERS (max %.3e)" % np.abs(oc - fc).max()))
            bad += not same
    # (even dimensions: the Fortran's bin index runs to xdim + 1 for x >= (xdim/2 - 1) binsize, so the rays stay below)
    for k3, (xd, yd, xr, yr) in enumerate(((10, 8, (-4., 3.), (-4., 2.)), (9, 7, (-4.4, 4.4), (-3.4, 3.4)))):
        bx, by = g.uniform(*xr, 600), g.uniform(*yr, 600)
        bl, bm = g.normal(0, 1e-3, 600), g.normal(0, 1e-3, 600)
        xa, ya, ph = (np.zeros((xd, yd), order="F") for _ in range(3))
        rec["southwellbin"](bx.copy(), by.copy(), bl.copy(), bm.copy(), 600, 1., xa, ya, ph, xd, yd)
        tag = "s%02d_southwellbin" % k3
        for nm, v in (("x", bx), ("y", by), ("l", bl), ("m", bm), ("xang_out", xa), ("yang_out", ya), ("phase_out", ph)):
            store[tag + "__" + nm] = v
        if check:
            o = of.reconstruct.southwellbin(bx, by, bl, bm, 1., xd, yd)
            same = all(np.array_equal(u, v, equal_nan=True) for u, v in zip(o, (xa, ya, ph)))
            print("%-44s %s" % (tag, "== oracle" if same else "DIFFERS"))
            bad += not same
    # specialFunctions.f95 called directly: Legendre polynomials and derivatives, radial Zernike polynomials, zernset
    sp = f95run.load(os.path.join(REF, "specialFunctions.f95"))
    xs = g.uniform(-1., 1., 6)
    leg = np.array([[[sp["legendre"](x, n), sp["legendrep"](x, n)] for x in xs] for n in range(9)], dtype=np.float64)
    rhos = np.concatenate([[0.], g.uniform(0., 1., 4)])
    nm = [(n, m) for n in range(9) for m in range(n % 2, n + 1, 2)]
    rad = np.array([[sp["radialpoly"](r_, n, m) for r_ in rhos] for n, m in nm], dtype=np.float64)
    ro, ao = np.array([n for n in range(5) for _ in range(n + 1)], dtype=np.int32), np.array([m for n in range(5) for m in range(-n, n + 1, 2)], dtype=np.int32)
    zs = []
    for r_, th in ((.3, .7), (0., .2), (.99, -2.)):
        po, dr, dt = (np.zeros(len(ro)) for _ in range(3))
        sp["zernset"](r_, th, ro, ao, len(ro), po, dr, dt)
        zs.append([po, dr, dt])
    store.update(sf_x=xs, sf_legendre=leg, sf_rho=rhos, sf_nm=np.array(nm), sf_radialpoly=rad, sf_rorder=ro, sf_aorder=ao,
                 sf_zernset_args=np.array([(.3, .7), (0., .2), (.99, -2.)]), sf_zernset=np.array(zs))
    if check:
        ok = all(of.specialfunctions.legendre(x, n) == leg[n, i, 0] and of.specialfunctions.legendrep(x, n) == leg[n, i, 1]
                 for n in range(9) for i, x in enumerate(xs))
        ok = ok and all(of.specialfunctions.radialpoly(r_, n, m) == rad[k, i] for k, (n, m) in enumerate(nm) for i, r_ in enumerate(rhos))
        for (r_, th), want in zip(((.3, .7), (0., .2), (.99, -2.)), zs):
            ok = ok and all(np.array_equal(u, v) for u, v in zip(of.specialfunctions.zernset(r_, th, ro, ao), want))
        print("%-44s %s" % ("specialFunctions (legendre, radialpoly, zernset)", "== oracle" if ok else "DIFFERS"))
        bad += not ok
    if "--no-write" not in sys.argv:
        np.savez_compressed(os.path.join(HERE, "f95_source.npz"), **store)
    print("%d cases, %d differ from the C oracle" % (len(store) and k + 1, bad) if check else "%d cases written" % (k + 1))


if __name__ == "__main__":
    main()
