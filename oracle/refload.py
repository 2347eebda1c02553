"""Import the reference's *unmodified* Python layer on top of the C oracle.

TEST INFRASTRUCTURE, BUILD-CONTAINER ONLY: ``/root/reference`` does not exist
on the GPU box, so nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py``
may call this.  It is used by ``tests/golden/make_golden.py`` (which writes the
committed fixtures) and by the ``not gpu`` tests that cross-check the oracle's
own numpy restatement against the real reference Python when it is present.

The reference imports several modules that are not installed here
(``matplotlib``, the author's ``utilities`` package, ``astropy``) and four
f2py extension modules that cannot be built (no Fortran compiler).  Stub
modules are injected for the former; the latter are filled with the oracle's
f2py-shaped namespaces.  The reference sources themselves are executed as they
lie on disk -- nothing is copied into this repository.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PXF_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "surfaces.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load(f2py_modules=None):
    """Return a namespace with the reference's sources/transformations/
    surfaces/analyses/conicsolve modules, bound to ``f2py_modules`` (default:
    the C oracle) for the four Fortran extension slots."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if f2py_modules is None:
        from . import f2py as _f
        f2py_modules = dict(transformationsf=_f.transformationsf, surfacesf=_f.surfacesf,
                            woltsurf=_f.woltsurf, zernsurf=_f.zernsurf,
                            specialfunctions=_f.specialfunctions)

    # --- third-party stubs (never exercised on the hot path) ---
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot")
        mpl.pyplot = plt
    util = _stub("utilities")
    img = _stub("utilities.imaging")
    util.imaging = img
    _noop = lambda *a, **k: None
    img.fitting = _stub("utilities.imaging.fitting", circle=_noop, circleMerit=_noop)
    img.analysis = _stub("utilities.imaging.analysis", ptov=_noop, rms=_noop)
    img.man = _stub("utilities.imaging.man")

    def _zmodes_unpinned(n):
        raise NotImplementedError("utilities.imaging.zernikemod.zmodes is an un-vendored third-party "
                                  "dependency of the reference; pass rorder/aorder explicitly")
    img.zernikemod = _stub("utilities.imaging.zernikemod", zmodes=_zmodes_unpinned)
    if "astropy" not in sys.modules:
        ap = _stub("astropy")
        ap.io = _stub("astropy.io")
        ap.io.fits = _stub("astropy.io.fits")

    # --- the package object, pointing at the read-only tree ---
    for k in [k for k in sys.modules if k == "PyXFocus" or k.startswith("PyXFocus.")]:
        del sys.modules[k]
    pkg = types.ModuleType("PyXFocus")
    pkg.__path__ = [REFERENCE_ROOT]
    sys.modules["PyXFocus"] = pkg

    # --- f2py slots ---
    for name in ("transformationsf", "surfacesf", "woltsurf", "zernsurf", "specialfunctions"):
        ns = f2py_modules[name]
        m = types.ModuleType("PyXFocus." + name)
        m.__dict__.update(vars(ns))
        sys.modules["PyXFocus." + name] = m
        setattr(pkg, name, m)
    _stub("PyXFocus.reconstruct")

    out = types.SimpleNamespace()
    for name in ("sources", "transformations", "conicsolve", "analyses", "surfaces"):
        setattr(out, name, importlib.import_module("PyXFocus." + name))
    return out
