"""The C-ABI library loads, exports every symbol include/pxf.h declares, and FAILS LOUDLY
without a GPU (no CPU fallback).  No compute is attempted here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pxf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pxf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pyxfocus_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 55
    for n in names:
        assert hasattr(L, n), "libpxf.so does not export %s" % n
    # the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().pxf_version() >= 100
    assert _lib.lib().pxf_newton_cap() == 1000


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "pxf.h")).read()
    for cite in ("transformationsf.f95:134-163", "surfacesf.f95:4-29", "woltsurf.f95:7-54", "woltsurf.f95:484-588",
                 "zernsurf.f95:8-101", "analyses.py", "transformations.py"):
        assert cite in text


def test_op_struct_layout_matches_header():
    from pyxfocus_b200 import _lib
    assert ctypes.sizeof(_lib.pxf_op) == 56          # int32 code, int32 reserved, double p[6]
    from pyxfocus_b200.program import OP, MAX_OPS
    text = open(os.path.join(ROOT, "include", "pxf.h")).read()
    for name, code in OP.items():
        m = re.search(r"PXF_OP_%s = (\d+)" % name, text)
        assert m and int(m.group(1)) == code, name
    assert int(re.search(r"#define PXF_MAX_OPS (\d+)", text).group(1)) == MAX_OPS


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import pyxfocus_b200 as pxf
    from pyxfocus_b200 import _lib
    x = np.zeros(8)
    ptr = x.ctypes.data
    rc = _lib.lib().pxf_reflect(ptr, ptr, ptr, ptr, ptr, ptr, 8, None, None)
    assert rc == 2                                   # PXF_ERR_CUDA
    assert b"no CUDA device" in _lib.lib().pxf_last_error()
    assert np.array_equal(x, np.zeros(8))            # nothing was computed on the host
    with pytest.raises(pxf.PxfError):
        pxf.sources.subannulus(1., 2., 1., 10)
    with pytest.raises(pxf.PxfError):
        pxf.transformationsf.reflect(*[np.zeros(4) for _ in range(6)])
    with pytest.raises(pxf.PxfError):
        pxf.host.trace([np.zeros(4) for _ in range(10)], pxf.Program().reflect())


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pyxfocus_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pyxfocus_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "pxfo_" not in src, f
