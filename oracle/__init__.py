"""CPU oracle for the PyXFocus hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this package.  ``pyxfocus_b200`` never does.

Parity status: the reference ships no tests or golden vectors and its Fortran cannot be compiled
in the build container (no Fortran compiler), so there is no ``oracle/_ref``.  The C restatement is
pinned against the Fortran SOURCE TEXT instead: ``oracle.f95run`` executes the ``.f95`` files
themselves (a translator with gfortran's arithmetic rules) and ``pxf_oracle.c`` reproduces every
routine's output bit for bit (``tests/test_f95_source.py``); see the header of ``pxf_oracle.c`` and
DESIGN.md.

``oracle.f2py`` exposes four namespaces named after the reference's f2py
extension modules (``transformationsf``, ``surfacesf``, ``woltsurf``,
``zernsurf``) with the positional signatures f2py generates from the .f95
sources, operating in place on numpy float64 arrays.
"""
from .f2py import transformationsf, surfacesf, woltsurf, zernsurf, specialfunctions, build, lib  # noqa: F401
