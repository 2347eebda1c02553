"""CPU oracle for the PyXFocus hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this package.  ``pyxfocus_b200`` never does.

Parity status: **unpinned** by the reference (it ships no tests or golden
vectors and its Fortran cannot be compiled in the build container); see the
header of ``pxf_oracle.c`` and DESIGN.md.

``oracle.f2py`` exposes four namespaces named after the reference's f2py
extension modules (``transformationsf``, ``surfacesf``, ``woltsurf``,
``zernsurf``) with the positional signatures f2py generates from the .f95
sources, operating in place on numpy float64 arrays.
"""
from .f2py import transformationsf, surfacesf, woltsurf, zernsurf, specialfunctions, build, lib  # noqa: F401
