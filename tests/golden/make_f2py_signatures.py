"""Pin the drop-in boundary mechanically (BUILD container only: needs /root/reference and numpy's f2py).

    python tests/golden/make_f2py_signatures.py      ->  tests/golden/f2py_signatures.json

1. ``f2py -h`` parses the reference's four Fortran sources (compiletrace.sh:1-4 builds exactly these into the
   extension modules ``transformationsf surfacesf woltsurf zernsurf``) into signature files; every subroutine is
   recorded with the argument order f2py gives the Python callable: required arguments in Fortran order, then the
   optional array-length arguments (``num``, ``arrsize``, ...), with each argument's type, rank and intent.
2. The reference's own wrappers (surfaces.py, transformations.py) are imported unmodified through oracle.refload on
   RECORDING stubs of the four modules and driven once each: which routine they call, with how many positional
   arguments of which kind -- the call sites the replacement modules must accept.

tests/test_f2py_boundary.py replays both against pyxfocus_b200.{transformationsf,surfacesf,woltsurf,zernsurf}.
"""
import json
import os
import re
import subprocess
import sys
import tempfile
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refload  # noqa: E402

MODULES = ("transformationsf", "surfacesf", "woltsurf", "zernsurf")


def parse_pyf(text):
    subs = {}
    cur = None
    for line in text.splitlines():
        line = line.strip()
        m = re.match(r"subroutine (\w+)\(([^)]*)\)", line)
        if m:
            cur = dict(fortran_order=[a.strip() for a in m.group(2).split(",") if a.strip()], args={})
            subs[m.group(1)] = cur
            continue
        if line.startswith("end subroutine"):
            cur = None
            continue
        if cur is not None and "::" in line:
            decl, names = line.split("::")
            name = names.split("=")[0].strip()
            typ = decl.split(",")[0].split(" ")[0].strip()
            dim = re.search(r"dimension\(([^)]*)\)", decl)
            intent = re.search(r"intent\((\w+)\)", decl)
            cur["args"][name] = dict(type=typ, rank=0 if not dim else dim.group(1).count(",") + 1,
                                     intent=intent.group(1) if intent else "in", optional="optional" in decl)
    out = {}
    for name, s in subs.items():
        req = [a for a in s["fortran_order"] if not s["args"][a]["optional"]]
        opt = [a for a in s["fortran_order"] if s["args"][a]["optional"]]
        out[name] = dict(python_order=req + opt, required=req, optional=opt, args=s["args"])
    return out


def kind(a):
    if isinstance(a, np.ndarray):
        return "int_array" if a.dtype.kind in "iu" else "array"
    if isinstance(a, (list, tuple)):
        return "sequence"
    return "scalar"


def record_calls():
    calls = []

    def recorder(module):
        class R:
            def __getattr__(self, name):
                def fn(*args, **kw):
                    calls.append(dict(module=module, routine=name, nargs=len(args), kinds=[kind(a) for a in args],
                                      kwargs=sorted(kw)))
                return fn
        return R()
    stubs = {m: recorder(m) for m in MODULES}
    stubs["specialfunctions"] = recorder("specialfunctions")
    mods = {k: SimpleNamespace(**{}) for k in stubs}
    # refload copies vars(ns) into module objects: hand it objects whose attribute lookup records
    ref = refload.load(f2py_modules={k: _AttrDict(v) for k, v in stubs.items()})
    src, tran, surf = ref.sources, ref.transformations, ref.surfaces
    np.random.seed(0)
    n = 8

    def rays():
        return src.subannulus(220., 221., .1, n, zhat=-1.)
    ind = np.arange(n) % 2 == 0
    ro, ao = np.array([0, 1, 1]), np.array([0, 1, -1])
    drive = [
        lambda: tran.transform(rays(), 1, 2, 3, .1, .2, .3), lambda: tran.transform(rays(), 1, 2, 3, .1, .2, .3, ind=ind),
        lambda: tran.itransform(rays(), 1, 2, 3, .1, .2, .3), lambda: tran.itransform(rays(), 1, 2, 3, .1, .2, .3, ind=ind),
        lambda: tran.reflect(rays()), lambda: tran.reflect(rays(), ind=ind), lambda: tran.refract(rays(), 1., 1.5),
        lambda: tran.radgrat(rays(), .01, 1, 2.4), lambda: tran.radgrat(rays(), .01, 1, np.full(n, 2.4)),
        lambda: tran.radgrat(rays(), .01, 1, 2.4, ind=ind), lambda: tran.grat(rays(), 160., 1, 2.4),
        lambda: surf.flat(rays()), lambda: surf.flat(rays(), nr=1.), lambda: surf.flat(rays(), ind=ind),
        lambda: surf.zernsurf(rays(), np.zeros(3), 10., rorder=ro, aorder=ao),
        lambda: surf.zernsurf(rays(), np.zeros(3), 10., rorder=ro, aorder=ao, nr=1.),
        lambda: surf.zernphase(rays(), np.zeros(3), 10., 1e-3, rorder=ro, aorder=ao),
        lambda: surf.sphere(rays(), 100.), lambda: surf.sphere(rays(), 100., nr=1.),
        lambda: surf.conic(rays(), 100., -1.), lambda: surf.conic(rays(), 100., -1., nr=1.),
        lambda: surf.cyl(rays(), 100.), lambda: surf.cyl(rays(), 100., nr=1.), lambda: surf.cylconic(rays(), 100., -1.),
        lambda: surf.torus(rays(), 100., 50.), lambda: surf.paraxial(rays(), 100.), lambda: surf.paraxialY(rays(), 100.),
        lambda: surf.wolterprimary(rays(), 220., 8400.), lambda: surf.wolterprimary(rays(), 220., 8400., nr=1.),
        lambda: surf.woltersecondary(rays(), 220., 8400.), lambda: surf.woltersine(rays(), 220., 8400., 1e-4, .1),
        lambda: surf.wsPrimary(rays(), 220., 8400., 1.), lambda: surf.wsSecondary(rays(), 220., 8400., 1.),
        lambda: surf.wsPrimaryB(rays(), 220., 8400., 1., .4), lambda: surf.wsSecondaryB(rays(), 220., 8400., 1., .4),
        lambda: surf.spoCone(rays(), 700., .01), lambda: surf.spoCone(rays(), 700., .01, ind=ind),
        lambda: surf.primaryLL(rays(), 220., 8400., 8500., 8400., .1, np.zeros(2), np.array([0, 1]), np.array([0, 0])),
        lambda: surf.secondaryLL(rays(), 220., 8400., 1., 8400., 8300., .1, np.zeros(2), np.array([0, 1]), np.array([0, 0])),
    ]
    failed = []
    for k, fn in enumerate(drive):
        try:
            fn()
        except Exception as e:       # noqa: BLE001  (a wrapper that is broken as shipped is recorded, not fatal)
            failed.append("%d: %s: %s" % (k, type(e).__name__, str(e)[:80]))
    return calls, failed


class _AttrDict(dict):
    """vars()-able view of a recorder: refload does ``m.__dict__.update(vars(ns))``; the module then gets a
    ``__getattr__`` that forwards every routine name to the recorder."""

    def __init__(self, rec):
        super().__init__()
        self["__getattr__"] = lambda name: getattr(rec, name)

    @property
    def __dict__(self):      # vars(obj) reads __dict__
        return dict(self)


def main():
    sigs = {}
    with tempfile.TemporaryDirectory() as tmp:
        for m in MODULES:
            pyf = os.path.join(tmp, m + ".pyf")
            subprocess.run(["f2py", "-h", pyf, "-m", m, os.path.join(refload.REFERENCE_ROOT, m + ".f95"), "--overwrite-signature"],
                           check=True, capture_output=True, cwd=tmp)
            sigs[m] = parse_pyf(open(pyf).read())
    calls, failed = record_calls()
    # de-duplicate call shapes
    seen, uniq = set(), []
    for c in calls:
        key = json.dumps(c, sort_keys=True)
        if key not in seen:
            seen.add(key)
            uniq.append(c)
    out = dict(source="f2py -h on /root/reference/{%s}.f95 (numpy %s); calls recorded from the reference's surfaces.py / "
                      "transformations.py through oracle.refload" % (",".join(MODULES), np.__version__),
               signatures=sigs, wrapper_calls=uniq, wrappers_broken_as_shipped=failed)
    with open(os.path.join(HERE, "f2py_signatures.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("routines:", {m: len(s) for m, s in sigs.items()}, "distinct wrapper calls:", len(uniq), "broken:", failed)


if __name__ == "__main__":
    main()
