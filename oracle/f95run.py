"""Run the reference's Fortran SOURCE TEXT without a Fortran compiler (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

There is no gfortran in this image, so ``oracle/pxf_oracle.c`` -- a hand restatement of the four ``.f95`` files -- could only
be pinned by known-answer tests.  This module is the second, mechanical reading: it translates the subset of Fortran 90
those files are written in into Python, statement by statement, and executes it with gfortran's arithmetic rules on
x86-64 (no ``-ffast-math``, no FMA):

* ``real*8`` is ``numpy.float64``, default ``real`` (and every literal without a ``d`` exponent, e.g. ``1.e-10`` or
  ``3.1415926535897931``) is ``numpy.float32``; mixed kinds promote exactly; an assignment rounds to the declared kind of its
  left-hand side; names that are not declared follow implicit typing (i-n integer, otherwise default real) unless the unit
  says ``implicit none``;
* integers are Python ints: ``/`` between two integers truncates toward zero, ``x**n`` with an integer ``n`` is repeated
  multiplication in ``__powidf2``'s order (square-and-multiply from the low bit), ``x**y`` with a real ``y`` is libm ``pow``;
* ``sin cos tan asin acos atan atan2 exp log pow`` are glibc's libm through ctypes -- the functions a gfortran build calls;
  ``sqrt`` and ``abs`` are IEEE operations;
* expressions are evaluated left to right in Fortran's precedence, arguments are passed by reference (a subroutine's
  changes to scalar or array-element arguments are written back at the call site), arrays are 1-based.

``load(path)`` returns ``{name: python callable}`` for every subroutine / function of a file (``include`` lines are
followed); a subroutine is called with its full Fortran argument list (``num`` included) and returns the tuple of its
arguments after the call.  ``tests/golden/make_f95_golden.py`` runs every routine of the hot path through this and
commits the vectors; ``tests/test_f95_source.py`` holds the C oracle to them bit for bit.
"""
import ctypes
import ctypes.util
import math
import os
import re

import numpy as np

f32, f64 = np.float32, np.float64

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("sin", "cos", "tan", "asin", "acos", "atan", "exp", "log"):
    for _name, _t in ((_n, ctypes.c_double), (_n + "f", ctypes.c_float)):
        _f = getattr(_libm, _name)
        _f.restype, _f.argtypes = _t, [_t]
for _n in ("atan2", "pow"):
    for _name, _t in ((_n, ctypes.c_double), (_n + "f", ctypes.c_float)):
        _f = getattr(_libm, _name)
        _f.restype, _f.argtypes = _t, [_t, _t]


def _real(x):
    """A Fortran real of the kind the value already has (integers become default real)."""
    return x if isinstance(x, (f32, f64)) else f32(x)


def _m1(name):
    d, s = getattr(_libm, name), getattr(_libm, name + "f")

    def fn(x):
        x = _real(x)
        return f32(s(float(x))) if isinstance(x, f32) else f64(d(float(x)))
    return fn


def _m2(name):
    d, s = getattr(_libm, name), getattr(_libm, name + "f")

    def fn(x, y):
        x, y = _real(x), _real(y)
        if isinstance(x, f32) and isinstance(y, f32):
            return f32(s(float(x), float(y)))
        return f64(d(float(x), float(y)))
    return fn


def _sqrt(x):
    x = _real(x)
    with np.errstate(all="ignore"):
        return np.sqrt(x)


def _abs(x):
    return abs(x)


def _sign(a, b):
    # sign(a, b): |a| with the sign of b
    if isinstance(a, int) and isinstance(b, int):
        return abs(a) if b >= 0 else -abs(a)
    r = type(_real(a))(math.copysign(float(abs(a)), float(b)))
    return r


def _mod(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return int(math.fmod(a, b))
    a, b = _real(a), _real(b)
    t = f64 if isinstance(a, f64) or isinstance(b, f64) else f32
    return t(math.fmod(float(a), float(b)))


def _int(x):
    return int(x) if isinstance(x, int) else int(math.trunc(float(x)))


def _floor(x):
    return int(math.floor(float(x)))


def _div(a, b):
    if isinstance(a, int) and isinstance(b, int):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    # an integer operand takes the kind of the real one
    if isinstance(a, int):
        a = type(b)(a)
    elif isinstance(b, int):
        b = type(a)(b)
    with np.errstate(all="ignore"):
        return a / b


def _powi(x, n):
    """libgcc __powidf2 / __powisf2: y = (n odd ? x : 1); while n >>= 1: x *= x; if n odd: y *= x; reciprocal for n < 0."""
    one = type(x)(1)
    m = abs(n)
    y = x if m & 1 else one
    with np.errstate(all="ignore"):
        m >>= 1
        while m:
            x = x * x
            if m & 1:
                y = y * x
            m >>= 1
        return one / y if n < 0 else y


def _pow(a, b):
    if isinstance(b, int):
        if isinstance(a, int):
            return a ** b if b >= 0 else (1 if a == 1 else (-1) ** b if a == -1 else 0)
        return _powi(a, b)
    a, b = _real(a), _real(b)
    if isinstance(a, f32) and isinstance(b, f32):
        return f32(_libm.powf(float(a), float(b)))
    return f64(_libm.pow(float(a), float(b)))


def _do(a, b, c=1):
    a, b, c = int(a), int(b), int(c)
    n = (b - a + c) // c if c > 0 else (a - b - c) // (-c)
    for k in range(max(n, 0)):
        yield a + k * c


INTRINSICS = {
    "sin": _m1("sin"), "cos": _m1("cos"), "tan": _m1("tan"), "asin": _m1("asin"), "acos": _m1("acos"), "atan": _m1("atan"),
    "exp": _m1("exp"), "log": _m1("log"), "atan2": _m2("atan2"), "sqrt": _sqrt, "abs": _abs, "dble": lambda x: f64(x),
    "real": lambda x: f32(x), "int": _int, "floor": _floor, "sign": _sign, "mod": _mod,
    "isnan": lambda x: bool(np.isnan(x)), "min": lambda *a: min(a), "max": lambda *a: max(a),
}

# ------------------------------------------------------------------------------------------------ source handling
_TOKEN = re.compile(r"""\s*(?:
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?)
  | (?P<dotop>\.(?:eq|ne|lt|le|gt|ge|and|or|not|true|false|eqv|neqv)\.)
  | (?P<name>[a-z_][a-z0-9_]*)
  | (?P<op>\*\*|==|/=|<=|>=|[-+*/(),<>=:])
)""", re.X)


def _tokens(text):
    out, pos = [], 0
    text = text.strip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise SyntaxError("cannot tokenise %r at %r" % (text, text[pos:pos + 10]))
        pos = m.end()
        kind = m.lastgroup
        out.append((kind, m.group(kind)))
    return out


def _logical_lines(path, seen=None):
    """Comment-free, lower-cased, continuation-joined lines of a file with its includes spliced in."""
    seen = seen or set()
    lines, buf = [], ""
    for raw in open(path, errors="replace").read().splitlines():
        line = raw
        m = re.match(r"\s*include\s+'([^']+)'", line, re.I)
        if m:
            inc = os.path.join(os.path.dirname(path), m.group(1))
            if inc not in seen:
                seen.add(inc)
                lines += _logical_lines(inc, seen)
            continue
        line = line.split("!")[0].rstrip().lower()
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith("&"):
            s = s[1:].strip()
        if buf:
            s = buf + " " + s
            buf = ""
        if s.endswith("&"):
            buf = s[:-1].rstrip()
            continue
        lines += [t.strip() for t in s.split(";") if t.strip()]
    return lines


def _logical_lines_no_include(path):
    """The file's own lines (its ``include`` lines dropped): which units a file DEFINES rather than pulls in."""
    import tempfile
    text = "\n".join(ln for ln in open(path, errors="replace").read().splitlines() if not re.match(r"\s*include\s+'", ln, re.I))
    with tempfile.NamedTemporaryFile("w", suffix=".f95", delete=False) as f:
        f.write(text)
    try:
        return _logical_lines(f.name)
    finally:
        os.unlink(f.name)


_TYPE = re.compile(r"^(real\*8|real\*4|double precision|real|integer|logical)\b(.*)$")
_UNIT = re.compile(r"^(?:recursive\s+)?(?:(real\*8|real\*4|real|integer|double precision)\s+)?(subroutine|function)\s+(\w+)\s*"
                   r"\(([^)]*)\)\s*(?:result\s*\((\w+)\))?$")


class _Unit:
    def __init__(self, kind, name, args, result, rtype):
        self.kind, self.name, self.args, self.result = kind, name, args, result or name
        self.types, self.dims, self.body, self.implicit_none = {}, {}, [], False
        if rtype:
            self.types[self.result] = rtype


def _split_top(text, sep=","):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def _parse_units(lines):
    units, cur = {}, None
    for ln in lines:
        m = _UNIT.match(ln)
        if m and cur is None:
            rtype, kind, name, args, result = m.groups()
            cur = _Unit(kind, name, [a.strip() for a in args.split(",") if a.strip()], result, rtype)
            continue
        if cur is None:
            continue
        if re.match(r"^end\s*(subroutine|function)?(\s+\w+)?$", ln) and not re.match(r"^end\s*(do|if)", ln):
            units[cur.name] = cur
            cur = None
            continue
        if ln == "implicit none":
            cur.implicit_none = True
            continue
        m = _TYPE.match(ln)
        if m and not re.match(r"^(real|integer)\s*\(", ln):
            typ, rest = m.group(1), m.group(2)
            attrs, _, ents = rest.partition("::")
            if not _:
                attrs, ents = "", rest
            dim = re.search(r"dimension\s*\(([^)]*(?:\([^)]*\))?[^)]*)\)", attrs)
            for ent in _split_top(ents):
                em = re.match(r"^(\w+)\s*(?:\((.*)\))?\s*(?:=\s*(.*))?$", ent)
                nm, d, init = em.group(1), em.group(2), em.group(3)
                cur.types[nm] = typ
                if d or dim:
                    cur.dims[nm] = _split_top(d if d else dim.group(1))
                if init:
                    cur.body.append("%s = %s" % (nm, init))
            continue
        cur.body.append(ln)
    return units


# ------------------------------------------------------------------------------------------------ expression translator
class _Expr:
    """Pratt parser over Fortran tokens; emits a Python expression string."""
    BIN = {".or.": (1, "or"), ".and.": (2, "and"), "==": (4, "=="), "/=": (4, "!="), "<": (4, "<"), "<=": (4, "<="),
           ">": (4, ">"), ">=": (4, ">="), ".eq.": (4, "=="), ".ne.": (4, "!="), ".lt.": (4, "<"), ".le.": (4, "<="),
           ".gt.": (4, ">"), ".ge.": (4, ">="), "+": (5, "+"), "-": (5, "-"), "*": (6, "*"), "/": (6, "/"), "**": (8, "**")}

    def __init__(self, toks, unit, units, lhs=False):
        self.t, self.i, self.u, self.units, self.lhs = toks, 0, unit, units, lhs

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def expr(self, minp=0):
        kind, val = self.peek()
        if val == ".not.":
            self.take()
            left = "(not %s)" % self.expr(3)
        elif val in ("-", "+"):
            self.take()
            # unary minus binds looser than ** and * /, tighter than binary + -
            left = "(%s%s)" % (val, self.expr(6))
        else:
            left = self.primary()
        while True:
            kind, val = self.peek()
            if val not in self.BIN:
                break
            prec, py = self.BIN[val]
            if prec < minp:
                break
            self.take()
            if val == "**":
                right = self.expr_unary_rhs(prec)             # right associative; -b allowed as exponent
                left = "_pow(%s, %s)" % (left, right)
            elif val == "/":
                right = self.expr(prec + 1)
                left = "_div(%s, %s)" % (left, right)
            else:
                right = self.expr(prec + 1)
                left = "(%s %s %s)" % (left, py, right)
        return left

    def expr_unary_rhs(self, prec):
        kind, val = self.peek()
        if val in ("-", "+"):
            self.take()
            return "(%s%s)" % (val, self.expr_unary_rhs(prec))
        return self.expr(prec)

    def primary(self):
        kind, val = self.take()
        if kind == "num":
            if re.fullmatch(r"\d+", val):
                return val
            if "d" in val:
                return "f64(%s)" % val.replace("d", "e")
            return "f32(%s)" % val
        if kind == "dotop":
            return {".true.": "True", ".false.": "False"}[val]
        if val == "(":
            e = self.expr()
            assert self.take()[1] == ")", "missing )"
            return "(%s)" % e
        if kind == "name":
            if self.peek()[1] == "(":
                self.take()
                args = []
                if self.peek()[1] != ")":
                    while True:
                        args.append(self.expr())
                        if self.peek()[1] == ",":
                            self.take()
                            continue
                        break
                assert self.take()[1] == ")", "missing ) after arguments of %s" % val
                if val in self.u.dims:
                    ref = "%s[%s]" % (val, ", ".join("(%s) - 1" % a for a in args))
                    if self.lhs:
                        return ref
                    return {"integer": "int(%s)", "real*8": "f64(%s)", "double precision": "f64(%s)"}.get(
                        self.u.types.get(val, "real"), "f32(%s)") % ref
                if val in self.units:
                    return "_U[%r](%s)" % (val, ", ".join(args))
                if val in INTRINSICS:
                    return "_I[%r](%s)" % (val, ", ".join(args))
                raise NameError("%s: %s(...) is neither an array, a unit of the file nor a known intrinsic" % (self.u.name, val))
            self.u_note(val)
            return val
        raise SyntaxError("unexpected token %r in %s" % (val, self.u.name))

    def u_note(self, name):
        if name not in self.u.types and name not in self.u.args:
            if self.u.implicit_none:
                raise NameError("%s: %s is not declared under implicit none" % (self.u.name, name))
            self.u.types[name] = "integer" if name[0] in "ijklmn" else "real"


def _tr(text, unit, units, lhs=False):
    p = _Expr(_tokens(text), unit, units, lhs)
    e = p.expr()
    if p.i != len(p.t):
        raise SyntaxError("trailing tokens in %r (%s)" % (text, unit.name))
    return e


def _cast(unit, name, expr):
    t = unit.types.get(name)
    if t is None:
        if unit.implicit_none:
            raise NameError("%s: %s is not declared under implicit none" % (unit.name, name))
        t = unit.types[name] = "integer" if name[0] in "ijklmn" else "real"
    return {"real*8": "f64(%s)", "double precision": "f64(%s)", "real*4": "f32(%s)", "real": "f32(%s)", "integer": "_I['int'](%s)",
            "logical": "bool(%s)"}[t] % expr


def _match_paren(s, start):
    depth = 0
    for k in range(start, len(s)):
        if s[k] == "(":
            depth += 1
        elif s[k] == ")":
            depth -= 1
            if depth == 0:
                return k
    raise SyntaxError("unbalanced parentheses in %r" % s)


def _statement(ln, unit, units, out, ind):
    """Translate one simple (non-block) statement; returns nothing, appends to out."""
    pad = "    " * ind
    if ln in ("exit",):
        out.append(pad + "break")
        return
    if ln == "cycle":
        out.append(pad + "continue")
        return
    if ln == "return":
        out.append(pad + "return _ret()")
        return
    if ln.startswith("print") or ln.startswith("write") or ln.startswith("deallocate"):
        out.append(pad + "pass")
        return
    m = re.match(r"^allocate\s*\((.*)\)$", ln)
    if m:
        for ent in _split_top(m.group(1)):
            em = re.match(r"^(\w+)\s*\((.*)\)$", ent)
            nm = em.group(1)
            dims = ", ".join("int(%s)" % _tr(d, unit, units) for d in _split_top(em.group(2)))
            dt = "np.float64" if unit.types.get(nm) in ("real*8", "double precision") else ("np.float32" if unit.types.get(nm, "real").startswith("real") else "np.int64")
            out.append(pad + "%s = np.zeros((%s,), dtype=%s)" % (nm, dims, dt))
        return
    m = re.match(r"^call\s+(\w+)\s*\((.*)\)$", ln)
    if m:
        name, args = m.group(1), _split_top(m.group(2))
        if name not in units:
            raise NameError("%s calls unknown subroutine %s" % (unit.name, name))
        exprs = [_tr(a, unit, units) for a in args]
        out.append(pad + "_t = _U[%r](%s)" % (name, ", ".join(exprs)))
        for k, a in enumerate(args):
            if re.fullmatch(r"\w+", a) and not a[0].isdigit() and a not in unit.dims:
                out.append(pad + "%s = _t[%d]" % (a, k))
            elif re.fullmatch(r"\w+\s*\(.*\)", a) and a.split("(")[0].strip() in unit.dims:
                nm = a.split("(")[0].strip()
                idx = ", ".join("(%s) - 1" % _tr(x, unit, units) for x in _split_top(a[a.index("(") + 1:-1]))
                out.append(pad + "%s[%s] = _t[%d]" % (nm, idx, k))
        return
    # assignment
    depth, eq = 0, -1
    for k, ch in enumerate(ln):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0 and ln[k - 1] not in "<>/=" and ln[k + 1:k + 2] != "=":
            eq = k
            break
    if eq < 0:
        raise SyntaxError("%s: cannot translate %r" % (unit.name, ln))
    lhs, rhs = ln[:eq].strip(), ln[eq + 1:].strip()
    r = _tr(rhs, unit, units)
    m = re.match(r"^(\w+)\s*\((.*)\)$", lhs)
    if m and m.group(1) in unit.dims:
        idx = ", ".join("(%s) - 1" % _tr(a, unit, units) for a in _split_top(m.group(2)))
        out.append(pad + "%s[%s] = %s" % (m.group(1), idx, r))         # the array's dtype rounds
    elif re.fullmatch(r"\w+", lhs):
        if lhs in unit.dims:
            out.append(pad + "%s[...] = %s" % (lhs, r))
        else:
            out.append(pad + "%s = %s" % (lhs, _cast(unit, lhs, r)))
    else:
        raise SyntaxError("%s: bad left-hand side %r" % (unit.name, lhs))


def _translate(unit, units):
    out = ["def %s(%s):" % (unit.name, ", ".join(unit.args))]
    body, ind = [], 1
    for ln in unit.body:
        pad = "    " * ind
        if re.match(r"^end\s*(do|if)$", ln):
            ind -= 1
            continue
        if re.match(r"^else\s*if\b", ln):
            k = ln.index("(")
            e = _match_paren(ln, k)
            body.append("    " * (ind - 1) + "elif %s:" % _tr(ln[k + 1:e], unit, units))
            continue
        if ln == "else":
            body.append("    " * (ind - 1) + "else:")
            continue
        if ln.startswith("if") and re.match(r"^if\s*\(", ln):
            k = ln.index("(")
            e = _match_paren(ln, k)
            cond, rest = _tr(ln[k + 1:e], unit, units), ln[e + 1:].strip()
            body.append(pad + "if %s:" % cond)
            if rest == "then":
                ind += 1
            else:
                _statement(rest, unit, units, body, ind + 1)
            continue
        m = re.match(r"^do\s+while\s*\((.*)\)$", ln)
        if m:
            body.append(pad + "while %s:" % _tr(m.group(1), unit, units))
            ind += 1
            continue
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", ln)
        if m:
            var, rng = m.group(1), _split_top(m.group(2))
            unit.types.setdefault(var, "integer")
            body.append(pad + "for %s in _do(%s):" % (var, ", ".join(_tr(r, unit, units) for r in rng)))
            body.append(pad + "    pass")
            ind += 1
            continue
        if ln == "do":
            body.append(pad + "while True:")
            ind += 1
            continue
        _statement(ln, unit, units, body, ind)
    # prologue: dummy scalars take their declared kind, locals start at zero, local arrays are allocated
    pro = []
    for a in unit.args:
        if a not in unit.dims:
            pro.append("    %s = %s" % (a, _cast(unit, a, a)))
    for nm, t in list(unit.types.items()):
        if nm in unit.args or nm in units or (nm == unit.name and unit.kind == "subroutine"):
            continue
        if nm in unit.dims:
            dims = unit.dims[nm]
            if any(d.strip() == ":" for d in dims):
                pro.append("    %s = None" % nm)
                continue
            dt = "np.float64" if t in ("real*8", "double precision") else ("np.float32" if t.startswith("real") else "np.int64")
            pro.append("    %s = np.zeros((%s,), dtype=%s)" % (nm, ", ".join("int(%s)" % _tr(d, unit, units) for d in dims), dt))
        else:
            pro.append("    %s = %s" % (nm, _cast(unit, nm, "0")))
    ret = "%s" % unit.result if unit.kind == "function" else "(%s,)" % ", ".join(unit.args)
    out.append("    def _ret():")
    out.append("        return %s" % ret)
    # _ret closes over the locals by name at call time: emit it as an inline return instead
    text = "\n".join(out[:1] + pro + body + ["    return %s" % ret])
    return text.replace("return _ret()", "return %s" % ret)


def load(path):
    """Translate a .f95 file (and its includes) and return {unit name: callable}."""
    units = _parse_units(_logical_lines(path))
    ns = {"np": np, "f32": f32, "f64": f64, "_I": INTRINSICS, "_pow": _pow, "_div": _div, "_do": _do, "_U": {}}
    src = {}
    for name, u in units.items():
        src[name] = _translate(u, units)
    for name, text in src.items():
        try:
            exec(compile(text, "<f95:%s>" % name, "exec"), ns)
        except SyntaxError as e:
            raise SyntaxError("%s\n---- translated %s ----\n%s" % (e, name, text))
        ns["_U"][name] = ns[name]
    ns["_U"]["__source__"] = src

    def guarded(fn):
        def call(*a):
            with np.errstate(all="ignore"):
                return fn(*a)
        return call
    return {k: (guarded(v) if callable(v) else v) for k, v in ns["_U"].items()}
