"""The drop-in boundary, pinned mechanically.  ``tests/golden/f2py_signatures.json`` holds (1) what ``f2py -h`` makes
of the reference's four Fortran sources -- for every subroutine the positional order of the Python callable, the
optional trailing length arguments, intents -- and (2) the calls the reference's own wrappers make (recorded from
surfaces.py / transformations.py on stub modules).  The replacement modules must present exactly those callables:
a renamed routine, a swapped argument or a missing optional fails here."""
import importlib
import inspect
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIN = json.load(open(os.path.join(ROOT, "tests", "golden", "f2py_signatures.json")))
# f2py exposes these helpers of the Fortran files too; no Python wrapper calls them (internal to the .f95)
INTERNAL = {"rotatevector", "rotateaxis", "zernset"}
# extensions the replacement adds after the f2py arguments (keyword use only)
EXTRA = {"mask"}


def routines():
    for mod, subs in PIN["signatures"].items():
        for name, sig in subs.items():
            if name not in INTERNAL:
                yield mod, name, sig


@pytest.mark.parametrize("mod,name,sig", list(routines()), ids=["%s.%s" % (m, n) for m, n, _ in routines()])
def test_every_f2py_routine_has_the_same_python_signature(mod, name, sig):
    m = importlib.import_module("pyxfocus_b200." + mod)
    assert hasattr(m, name), "pyxfocus_b200.%s lacks %s" % (mod, name)
    params = list(inspect.signature(getattr(m, name)).parameters.values())
    names = [p.name for p in params if p.name not in EXTRA]
    assert names == sig["python_order"], "%s.%s: %s != f2py's %s" % (mod, name, names, sig["python_order"])
    for p in params:
        if p.name in sig["required"]:
            assert p.default is inspect.Parameter.empty, "%s.%s: %s must be required" % (mod, name, p.name)
        else:
            assert p.default is None, "%s.%s: %s must be optional (f2py: =shape(...))" % (mod, name, p.name)


def test_no_routine_is_missing_or_invented():
    for mod, subs in PIN["signatures"].items():
        m = importlib.import_module("pyxfocus_b200." + mod)
        public = {n for n, f in vars(m).items() if inspect.isfunction(f) and not n.startswith("_") and f.__module__ == m.__name__}
        want = set(subs) - INTERNAL
        assert want <= public, "%s lacks %s" % (mod, sorted(want - public))
        assert public <= want, "%s has routines the reference's module does not: %s" % (mod, sorted(public - want))


def test_wrapper_calls_bind_to_the_replacement_signatures():
    """Every call shape the reference's wrappers make (positional arity) binds to the replacement callable, and the
    argument kinds land on parameters of the right nature (arrays on array parameters, scalars on scalars)."""
    assert len(PIN["wrapper_calls"]) >= 30
    for c in PIN["wrapper_calls"]:
        m = importlib.import_module("pyxfocus_b200." + c["module"])
        fn = getattr(m, c["routine"])
        sig = PIN["signatures"][c["module"]][c["routine"]]
        bound = inspect.signature(fn).bind(*([None] * c["nargs"]))       # raises TypeError on an arity mismatch
        assert len(bound.arguments) == c["nargs"]
        assert c["nargs"] == len(sig["required"]), (c, sig["required"])
        for kind, arg in zip(c["kinds"], sig["python_order"]):
            rank = sig["args"][arg]["rank"]
            if c["routine"] in ("grat",) and arg in ("order", "wave"):
                continue        # transformations.grat passes its order/wave through: arrays in the Fortran, so callers must pass rows
            assert (kind in ("array", "int_array", "sequence")) == (rank > 0), (c["routine"], arg, kind, rank)


@pytest.mark.gpu
def test_recorded_wrapper_calls_run_on_the_device():
    """Replay: each recorded call shape, with device rows / numpy tables / floats of the recorded kinds, executes on
    the replacement module (in-place on the rows, None returned)."""
    import torch
    import pyxfocus_b200 as pxf
    n = 64
    rng = np.random.default_rng(0)
    for c in PIN["wrapper_calls"]:
        m = importlib.import_module("pyxfocus_b200." + c["module"])
        sig = PIN["signatures"][c["module"]][c["routine"]]
        rays = pxf.sources.subannulus(220., 221., .1, n, zhat=-1., rng="philox", seed=1, device="cuda")
        pxf.transformations.transform(rays, 0, 0, -8400., 0, 0, 0)
        pxf.surfaces.wolterprimary(rays, 220., 8400.)            # gives the bundle surface normals (reflect / refract)
        rows = dict(zip(["opd", "x", "y", "z", "l", "m", "n", "ux", "uy", "uz"], rays))
        args = []
        ntab = 3
        for kind, arg in zip(c["kinds"], sig["python_order"]):
            if arg in rows and kind == "array":
                args.append(rows[arg])
            elif kind == "array":
                per_ray = sig["args"][arg]["rank"] > 0 and arg in ("wave",)
                args.append(torch.full((n,), 2.4e-6, dtype=torch.float64, device="cuda") if per_ray else np.full(ntab, 1e-4))
            elif kind == "int_array":
                args.append(np.array([0, 1, -1][:ntab]) if arg.startswith("aorder") else np.array([0, 1, 1][:ntab]))   # Zernike azimuthal orders may be negative (sine terms); Legendre orders not
            else:
                args.append({"r0": 220., "z0": 8400., "psi": 1., "rad": 100., "r": 1000., "k": -1., "n1": 1., "n2": 1.5,
                             "alpha": .0065, "tg": .0065, "zmax": 8500., "zmin": 8400., "dphi": .1, "nr": 1.,
                             "d": 160., "dpermm": .01, "order": 1., "wave": 2.4e-6, "thick": .4, "f": 100.,
                             "rin": 100., "rout": 50., "amp": 1e-4, "freq": .1}.get(arg, .01))
        before = [r.clone() for r in rays]
        # a recorded SCALAR where the Fortran declares an intent(inout) array (transformations.grat's order, wave): f2py
        # refuses a Python float there unless num == 1, and so does the replacement; the call a user can make passes rows
        lifted = [k for k, (kind, arg) in enumerate(zip(c["kinds"], sig["python_order"]))
                  if kind == "scalar" and sig["args"][arg]["rank"] > 0]
        if lifted:
            with pytest.raises(ValueError):
                getattr(m, c["routine"])(*args)
            for k in lifted:
                args[k] = torch.full((n,), float(args[k]), dtype=torch.float64, device="cuda")
        out = getattr(m, c["routine"])(*args)
        assert out is None, c["routine"]
        torch.cuda.synchronize()
        changed = any(not torch.equal(a, b) for a, b in zip(before, rays))
        assert changed, "%s.%s left every row untouched" % (c["module"], c["routine"])
