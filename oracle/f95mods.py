"""f2py-shaped modules whose routines are the reference's Fortran SOURCE TEXT executed by ``oracle.f95run``
(TEST INFRASTRUCTURE ONLY).  ``modules()`` is a drop-in for the ``f2py_modules`` argument of ``oracle.refload.load``:
the reference's unmodified Python layer then runs over the reference's unmodified Fortran -- the whole reference,
executed in a container that has no Fortran compiler (slowly: use a few hundred rays).

f2py's wrapper conventions restated: array-length arguments (``num``, ``arrsize``, ``cnum`` ...) are hidden and derived
from the arrays; ``intent(inout)`` arrays are mutated in place; every routine returns ``None``."""
import os
from types import SimpleNamespace

import numpy as np

from . import f95run
from .refload import REFERENCE_ROOT

# hidden dimension argument -> the array whose length it is (the f2py signature files say so: tests/golden/f2py_signatures.json)
_LENGTH_OF = {"num": ("x", "l", "opd"), "arrsize": ("coeff",), "arrsize1": ("coeff1",), "arrsize2": ("coeff2",), "cnum": ("coeff",),
              "nc": ("coeff",), "np": ("p",), "znum": ("rorder",)}


def _wrap(fn, fargs):
    visible = [a for a in fargs if a not in _LENGTH_OF]

    def call(*args):
        if len(args) != len(visible):
            raise TypeError("expected %d arguments (%s), got %d" % (len(visible), ", ".join(visible), len(args)))
        given = dict(zip(visible, args))
        full = []
        for a in fargs:
            if a in given:
                v = given[a]
                if isinstance(v, np.ndarray) and v.dtype.kind == "f" and v.dtype != np.float64:
                    raise TypeError("intent(inout) array must be float64")
                full.append(v)
            else:
                src = next(n for n in _LENGTH_OF[a] if n in given)
                full.append(int(np.size(given[src])))
        fn(*full)
        return None
    return call


def _module(name):
    path = os.path.join(REFERENCE_ROOT, name + ".f95")
    units = f95run._parse_units(f95run._logical_lines(path))
    fns = f95run.load(path)
    own = {u.name for u in f95run._parse_units([ln for ln in f95run._logical_lines_no_include(path)]).values()}
    return SimpleNamespace(**{n: _wrap(fns[n], units[n].args) for n in own if units[n].kind == "subroutine"})


def modules():
    mods = {n: _module(n) for n in ("transformationsf", "surfacesf", "woltsurf", "zernsurf")}
    sp = f95run.load(os.path.join(REFERENCE_ROOT, "specialFunctions.f95"))
    mods["specialfunctions"] = SimpleNamespace(**{k: v for k, v in sp.items() if not k.startswith("__")})
    return mods
