"""Host-side logic that needs no GPU: program recording, mask conversion, sharding, 4x4
coordinate bookkeeping, argument validation."""
import numpy as np
import pytest
import torch


def test_program_records_reference_calls_with_negated_transform_args():
    import pyxfocus_b200 as pxf
    from pyxfocus_b200.program import OP, recorder_for
    rays = [torch.zeros(4, dtype=torch.float64) for _ in range(10)]     # never executed
    with pytest.raises(Exception):
        with pxf.fused(rays):
            pxf.transformations.transform(rays, 1., 2., 3., .1, .2, .3)
            pxf.surfaces.wolterprimary(rays, 220., 8400.)
            pxf.transformations.reflect(rays)
            pxf.surfaces.wsPrimary(rays, 220., 1.e4, 1.)
            pxf.surfaces.flat(rays, nr=1.5)
            prog = recorder_for(rays)
            assert [c for c, _ in prog.ops] == [OP["TRANSFORM"], OP["WOLTERPRIMARY"], OP["REFLECT"],
                                                OP["WSPRIMARY"], OP["FLATOPD"]]
            assert prog.ops[0][1] == [-1., -2., -3., -.1, -.2, -.3]     # transformations.py:29
            assert prog.ops[1][1] == [220., 8400., 1.]
            assert prog.ops[3][1][0] == pytest.approx(.25 * np.arctan(220. / 1.e4))   # alpha from woltparam
            raise KeyError("stop before the (GPU) flush")
    assert recorder_for(rays) is None                                    # context cleaned up


def test_program_limits():
    import pyxfocus_b200 as pxf
    p = pxf.Program()
    with pytest.raises(ValueError):
        p.add(1, *range(7))
    p.vignette_box(3, 1., 2.)
    assert p.has_vignette() and len(p) == 1
    ops = p.c_ops()
    assert ops[0].code == 18 and list(ops[0].p)[:3] == [3., 1., 2.]


def test_mask_conversion():
    from pyxfocus_b200._call import to_mask
    cpu = torch.device("cpu")
    m = np.array([True, False, True, False, False])
    assert to_mask(m, 5, cpu).tolist() == [1, 0, 1, 0, 0]
    assert to_mask(np.where(m), 5, cpu).tolist() == [1, 0, 1, 0, 0]          # np.where tuple
    assert to_mask(np.array([4, 4, 0]), 5, cpu).tolist() == [1, 0, 0, 0, 1]  # index array with repeats
    assert to_mask(torch.tensor([False, True, True, False, False]), 5, cpu).tolist() == [0, 1, 1, 0, 0]
    assert to_mask(np.array([], dtype=np.int64), 5, cpu).tolist() == [0] * 5
    with pytest.raises(IndexError):
        to_mask(np.array([True, False]), 5, cpu)


def test_shard_ranges_cover_the_bundle_in_rank_order():
    from pyxfocus_b200.dist import shard_range
    for num in (0, 1, 7, 1000, 10 ** 9 + 3):
        for world in (1, 2, 4, 8):
            edges = [shard_range(num, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == num
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_coordinate_bookkeeping_roundtrip():
    from pyxfocus_b200 import transformations as T
    c = T.newCoords()
    T._update_coords_fwd(c, 1., -2., 3., .1, .2, -.3)
    assert np.allclose(c[1] @ c[3], np.identity(4), atol=1e-14)      # global->local then local->global
    assert np.allclose(c[0] @ c[2], np.identity(4), atol=1e-14)
    assert np.allclose(T.rotationM(.1, .2, .3) @ T.rotationM(.1, .2, .3, inverse=True), np.identity(4), atol=1e-15)
    # a point at the new origin maps to zero
    p = c[1] @ np.array([1., -2., 3., 1.])
    assert np.allclose(p[:3], 0., atol=1e-14)


def test_bundle_alloc_alignment_rule():
    from pyxfocus_b200._call import bundle_alloc
    for n in (0, 1, 2, 5, 1000, 1001):
        rows = bundle_alloc(n, "cpu")
        assert len(rows) == 10 and all(r.shape[0] == n and r.is_contiguous() for r in rows)
        # consecutive rows are an even number of doubles apart -> every row keeps 16-byte alignment
        if n:
            assert ((rows[1].data_ptr() - rows[0].data_ptr()) // 8) % 2 == 0


def test_host_trace_argument_validation():
    import pyxfocus_b200 as pxf
    rows = [np.zeros(4) for _ in range(10)]
    with pytest.raises(ValueError):
        pxf.host.trace(rows, pxf.Program())                              # empty program
    rows[3] = np.zeros(5)
    with pytest.raises(ValueError):
        pxf.host.trace(rows, pxf.Program().reflect())                    # ragged rows
    rows[3] = np.zeros(4, dtype=np.float32)
    with pytest.raises(ValueError):
        pxf.host.trace(rows, pxf.Program().reflect())                    # wrong dtype


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract: ONE JSON line on stdout (the driver parses it); everything else -- including what
    libraries write to fd 1 -- goes to stderr.  The reference arm runs on the CPU, so it is checked here."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-rays", "2e5"], capture_output=True, text=True, timeout=300, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_segmented_program_host_side_validation():
    """SegmentedProgram folds its tables on the host (no GPU needed up to the upload): shape checks and the
    C-side fill (same routine sequence in every segment, seg_start monotone)."""
    import ctypes
    import pyxfocus_b200 as pxf
    from pyxfocus_b200 import _lib
    L = _lib.lib()
    ops = (_lib.pxf_op * 4)()
    for k, (code, p) in enumerate([(10, (220., 8400., 1.)), (3, ()), (10, (300., 8000., 1.)), (3, ())]):
        ops[k].code = code
        for j, v in enumerate(p):
            ops[k].p[j] = v
    start = np.array([0, 10, 25], dtype=np.int64)
    nbytes = int(L.pxf_segmented_table_bytes(2, 2))
    assert nbytes > 0
    buf = np.zeros(nbytes, dtype=np.uint8)
    assert L.pxf_segmented_table_fill(ops, 2, 2, start.ctypes.data, buf.ctypes.data) == 0
    hdr = buf[:16].view(np.int32)
    assert hdr[2] == 2 and hdr[3] == 2                      # nops, nseg
    ops[3].code = 6                                         # second segment runs a different routine
    assert L.pxf_segmented_table_fill(ops, 2, 2, start.ctypes.data, buf.ctypes.data) != 0
    assert b"same opcode sequence" in L.pxf_last_error()
    ops[3].code = 3
    bad = np.array([0, 30, 25], dtype=np.int64)
    assert L.pxf_segmented_table_fill(ops, 2, 2, bad.ctypes.data, buf.ctypes.data) != 0
    with pytest.raises(ValueError):
        pxf.SegmentedProgram([pxf.Program().flat()], [3, 4])
