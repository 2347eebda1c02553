"""Fused per-ray programs: a whole chain of f2py routines in ONE kernel launch.

The reference runs a trace as a sequence of Fortran calls, each a full pass over the
bundle in memory (~7 passes for a Wolter-I pair).  ``Program`` records the
same routines -- same names and argument meaning as the f2py modules -- and
``pxf_trace_program`` executes the list per ray in registers: rows are read from HBM once
and written once.  The arithmetic is the same device code as the per-routine kernels, so the
result is bit-identical to issuing the routines one by one.

Two ways to use it:

* explicitly::

      prog = Program().transform(0, 0, 8400., 0, 0, 0).wolterprimary(220., 8400., 1.).reflect()
      prog.run(rays)

* transparently, with the reference's own call sequence::

      with fused(rays):
          tran.transform(rays, 0, 0, -8400., 0, 0, 0)
          surf.wolterprimary(rays, 220., 8400.)
          tran.reflect(rays)
      # one kernel launch happens on exit (or as soon as something needs the rays)
"""
import ctypes

import torch

from . import _lib
from ._call import stream_ptr

OP = dict(TRANSFORM=1, ITRANSFORM=2, REFLECT=3, REFRACT=4, RADGRAT=5, FLAT=6, FLATOPD=7, CONIC=8,
          CONICOPD=9, WOLTERPRIMARY=10, WOLTERPRIMARYOPD=11, WOLTERSECONDARY=12, WOLTERSINE=13,
          WSPRIMARY=14, WSSECONDARY=15, SPOCONE=16, VIGNETTE_MAG=17, VIGNETTE_BOX=18,
          VIGNETTE_ABS=19, KICK=20, ZERNSURF=21, KICKN=22, VIGNETTE_RHOGT=23, GRATFAN=24, ROTX_REMAINING=25)
MAX_OPS = 24
_VIGNETTES = (OP["VIGNETTE_MAG"], OP["VIGNETTE_BOX"], OP["VIGNETTE_ABS"], OP["VIGNETTE_RHOGT"])
_AUX_OPS = (OP["GRATFAN"], OP["ROTX_REMAINING"])


class Program:
    """An ordered list of per-ray operations (arguments as the Fortran routine sees them)."""

    def __init__(self):
        self.ops = []
        self._tables = {}        # op index -> host table (uint8 tensor) of a ZERNSURF op
        self._dev_tables = {}    # (op index, device) -> its device copy

    def __len__(self):
        return len(self.ops)

    def add(self, code, *p):
        if len(p) > 6:
            raise ValueError("an op carries at most 6 scalars")
        self.ops.append((int(code), [float(v) for v in p]))
        return self

    # --- transformationsf ---
    def transform(self, tx, ty, tz, rx, ry, rz):
        return self.add(OP["TRANSFORM"], tx, ty, tz, rx, ry, rz)

    def itransform(self, tx, ty, tz, rx, ry, rz):
        return self.add(OP["ITRANSFORM"], tx, ty, tz, rx, ry, rz)

    def reflect(self):
        return self.add(OP["REFLECT"])

    def refract(self, n1, n2):
        return self.add(OP["REFRACT"], n1, n2)

    def radgrat(self, wave, dpermm, order):
        return self.add(OP["RADGRAT"], wave, dpermm, order)

    # --- surfacesf ---
    def flat(self):
        return self.add(OP["FLAT"])

    def flatopd(self, nr):
        return self.add(OP["FLATOPD"], nr)

    def conic(self, r, k):
        return self.add(OP["CONIC"], r, k)

    def conicopd(self, r, k, nr):
        return self.add(OP["CONICOPD"], r, k, nr)

    # --- woltsurf ---
    def wolterprimary(self, r0, z0, psi):
        return self.add(OP["WOLTERPRIMARY"], r0, z0, psi)

    def wolterprimaryopd(self, r0, z0, psi, nr):
        return self.add(OP["WOLTERPRIMARYOPD"], r0, z0, psi, nr)

    def woltersecondary(self, r0, z0, psi):
        return self.add(OP["WOLTERSECONDARY"], r0, z0, psi)

    def woltersine(self, r0, z0, amp, freq):
        return self.add(OP["WOLTERSINE"], r0, z0, amp, freq)

    def wsprimary(self, alpha, z0, psi):
        return self.add(OP["WSPRIMARY"], alpha, z0, psi)

    def wssecondary(self, alpha, z0, psi):
        return self.add(OP["WSSECONDARY"], alpha, z0, psi)

    def spocone(self, r0, tg):
        return self.add(OP["SPOCONE"], r0, tg)

    # --- zernsurf ---
    def zernsurf(self, coeff, rorder, aorder, rad, nr=None):
        """tracezern / tracezernOPD (zernsurf.f95:8-203) inside the program; radial orders <= 7, one Zernike surface
        per program.  The term list is folded on the host exactly as for the stand-alone routine; the table is
        uploaded once per device and staged in shared memory by the kernel."""
        import numpy as np
        c = np.ascontiguousarray(coeff, dtype=np.float64)
        r = np.ascontiguousarray(rorder, dtype=np.int32)
        a = np.ascontiguousarray(aorder, dtype=np.int32)
        if not (c.shape == r.shape == a.shape) or c.ndim != 1 or c.shape[0] == 0:
            raise ValueError("coeff, rorder, aorder must be 1-D arrays of one length")
        if any(code == OP["ZERNSURF"] for code, _ in self.ops):
            raise ValueError("a program carries at most one Zernike surface")
        L = _lib.lib()
        tab = torch.empty(int(L.pxf_zern_table_bytes()), dtype=torch.uint8)
        nmax = int(L.pxf_zern_table_fill(c.ctypes.data, r.ctypes.data, a.ctypes.data, c.shape[0], float(rad),
                                         0 if nr is None else 1, 0. if nr is None else float(nr), tab.data_ptr()))
        if nmax < 0:
            raise ValueError("invalid Zernike term list")
        if nmax > 7:
            raise NotImplementedError("fused Zernike surfaces support radial orders <= 7")
        self._tables[len(self.ops)] = tab
        return self.add(OP["ZERNSURF"], 0., 0. if nr is None else 1., float(nmax))

    # --- per-ray predicates (the ray stops at the op; use with ``alive``) ---
    def vignette_mag(self):
        """keep rays with l^2+m^2+n^2 > .1 (transformations.py:220-223)"""
        return self.add(OP["VIGNETTE_MAG"])

    def vignette_box(self, row, lo, hi):
        """keep rays with lo < rays[row] < hi"""
        return self.add(OP["VIGNETTE_BOX"], row, lo, hi)

    def vignette_abs(self, row, hi, centre=0.):
        """keep rays with |rays[row] - centre| < hi"""
        return self.add(OP["VIGNETTE_ABS"], row, hi, centre)

    def vignette_rhogt(self, rho0):
        """keep rays with sqrt(x^2+y^2) > rho0 (examples/axro/axialHeights.py:285-286)"""
        return self.add(OP["VIGNETTE_RHOGT"], rho0)

    def kickn(self, dl, dm):
        """l += dl, m += dm, n = -sqrt(n^2-dl^2-dm^2) (pointing offsets, examples/arcus/cat.py:246-249)"""
        return self.add(OP["KICKN"], dl, dm)

    def gratfan(self, ang, hubdist, l, dpermm, order, wave):
        """The fanned radial-grating array of examples/arcus/sector.py:681-707 as one per-ray loop: each ray is
        rotated about the hub axis by ``ang`` per grating (traced to the grating plane when steep enough) until it
        lands between ``hubdist`` and ``l + hubdist`` from the hub, then reflected and diffracted.  ``wave``: a
        float [nm] (radgrat) or a per-ray CUDA tensor (radgratW).  ``run`` needs ``aux=FanAux(...)``: the per-ray
        grating index and its maximum come back there."""
        if isinstance(wave, torch.Tensor):
            self._wave = wave
            w = float("nan")
        else:
            w = float(wave)
        return self.add(OP["GRATFAN"], ang, hubdist, l, dpermm, order, w)

    def rotx_remaining(self, ang, total):
        """transform(0,0,0,ang,0,0) applied (total - count[i]) times: the whole-bundle fan rotations a ray still
        receives after it met its grating (sector.py:693), with ``count`` from a previous ``gratfan``."""
        return self.add(OP["ROTX_REMAINING"], ang, total)

    def kick(self, dl, dm, sn):
        """l += dl, m += dm, n = sn*sqrt(1-l^2-m^2) (field-angle kick, axialHeights.py:94-95)"""
        return self.add(OP["KICK"], dl, dm, sn)

    # --- execution ---
    def has_vignette(self):
        return any(c in _VIGNETTES for c, _ in self.ops)

    def c_ops(self, device=None):
        import struct
        arr = (_lib.pxf_op * len(self.ops))()
        for k, (code, p) in enumerate(self.ops):
            arr[k].code = code
            for j, v in enumerate(p):
                arr[k].p[j] = v
            if k in self._tables:
                if device is None:
                    raise ValueError("a program with a Zernike surface needs a device")
                key = (k, str(device))
                if key not in self._dev_tables:
                    self._dev_tables[key] = self._tables[k].to(device)
                # the table's device address travels bit-cast in p[0] (include/pxf.h, PXF_OP_ZERNSURF)
                arr[k].p[0] = struct.unpack("d", struct.pack("Q", self._dev_tables[key].data_ptr()))[0]
        return arr

    def precompile(self, segmented=False):
        """Compile the run-time specialised kernels this program will want (no GPU needed; cached on disk under
        ``pyxfocus_b200/_jit``).  Returns the number of kernels available (0: NVRTC unavailable -- the interpreter
        will be used)."""
        import struct
        arr = (_lib.pxf_op * len(self.ops))()
        for k, (code, p) in enumerate(self.ops):
            arr[k].code = code
            for j, v in enumerate(p):
                arr[k].p[j] = v
            if k in self._tables:
                arr[k].p[0] = struct.unpack("d", struct.pack("Q", 8))[0]      # any non-null table address
        n = int(_lib.lib().pxf_jit_compile(arr, len(self.ops), 1 if segmented else 0, None))
        if n < 0:
            _lib.check(1)
        return n

    def run(self, rays, alive=None, out=None, sums=None, aux=None):
        """Execute on a bundle (list of ten CUDA fp64 rows).  Returns the ``alive`` uint8
        tensor when the program contains a vignette predicate (allocated if not given).
        ``out``: optional second bundle -- rows are read from ``rays`` and every row the program
        touches is written to ``out``; ``rays`` is left untouched.
        ``sums``: optional CUDA float64 tensor (>= 16 entries) that receives, from the same
        kernel, {count, sum x, sum y, count} of the final bundle over the surviving rays --
        pass it on to ``analyses.hpd(..., sums=sums)`` to skip the centroid pass."""
        if not self.ops:
            return alive
        if len(self.ops) > MAX_OPS:
            # split into several launches; still one HBM round trip per <=24 elements
            head, tail = Program(), Program()
            head.ops, tail.ops = self.ops[:MAX_OPS], self.ops[MAX_OPS:]
            head._tables = {k: v for k, v in self._tables.items() if k < MAX_OPS}
            tail._tables = {k - MAX_OPS: v for k, v in self._tables.items() if k >= MAX_OPS}
            if head.has_vignette() or tail.has_vignette():
                raise ValueError("programs with vignette predicates are limited to %d ops" % MAX_OPS)
            if out is not None:
                # the head stores only the rows it touches: give `out` every row first, then run both parts in place
                for k in range(10):
                    if rays[k] is not None and out[k] is not None:
                        out[k].copy_(rays[k])
                rays = out
            head.run(rays)
            return tail.run(rays, sums=sums)
        dev = rays[1].device
        num = rays[1].shape[0]
        for r in rays:
            if r is not None and (not r.is_cuda or r.dtype != torch.float64 or not r.is_contiguous()
                                  or r.shape[0] != num):
                raise ValueError("ray rows must be contiguous 1-D float64 CUDA tensors of equal length")
        if self.has_vignette() and alive is None:
            alive = torch.empty(num, dtype=torch.uint8, device=dev)
        ptrs = (ctypes.c_void_p * 10)(*[(r.data_ptr() if r is not None else None) for r in rays])
        ops = self.c_ops(dev)
        ap = alive.data_ptr() if alive is not None else None
        optrs = None
        if out is not None:
            optrs = (ctypes.c_void_p * 10)(*[(r.data_ptr() if r is not None else None) for r in out])
        uses_aux = any(c in _AUX_OPS for c, _ in self.ops)
        if uses_aux and aux is None:
            raise ValueError("programs with gratfan / rotx_remaining need aux=FanAux(...)")
        with torch.cuda.device(dev):
            L = _lib.lib()
            if uses_aux:
                if sums is not None and (not sums.is_cuda or sums.dtype != torch.float64 or sums.numel() < 16):
                    raise ValueError("sums must be a contiguous CUDA float64 tensor with >= 16 entries")
                scratch = (torch.empty(int(L.pxf_sums_scratch_bytes()), dtype=torch.uint8, device=dev)
                           if sums is not None else None)
                wave = getattr(self, "_wave", None)
                if wave is not None and (not wave.is_cuda or wave.dtype != torch.float64 or not wave.is_contiguous()
                                         or wave.shape[0] != num):
                    raise ValueError("wave must be a contiguous CUDA float64 tensor with one entry per ray")
                ca = aux.c_struct(num, dev, wave)
                rc = L.pxf_trace_program_aux(ptrs, optrs, num, ops, len(self.ops), ap, ctypes.byref(ca),
                                             sums.data_ptr() if sums is not None else None,
                                             scratch.data_ptr() if scratch is not None else None, stream_ptr(dev))
            elif sums is not None:
                if not sums.is_cuda or sums.dtype != torch.float64 or sums.numel() < 16 or not sums.is_contiguous():
                    raise ValueError("sums must be a contiguous CUDA float64 tensor with >= 16 entries")
                scratch = torch.empty(int(L.pxf_sums_scratch_bytes()), dtype=torch.uint8, device=dev)
                rc = L.pxf_trace_program_sums(ptrs, optrs, num, ops, len(self.ops), ap, sums.data_ptr(),
                                              scratch.data_ptr(), stream_ptr(dev))
            elif out is None:
                rc = L.pxf_trace_program(ptrs, num, ops, len(self.ops), ap, stream_ptr(dev))
            else:
                rc = L.pxf_trace_program_to(ptrs, optrs, num, ops, len(self.ops), ap, stream_ptr(dev))
        _lib.check(rc)
        return alive


class FanAux:
    """Side arrays of a grating-fan program (include/pxf.h, pxf_program_aux): ``count`` int32[num] = gratings passed
    per ray (-1: none met), ``count_max`` int32[1] = the largest (cap + 1 if some ray met none)."""

    def __init__(self, num, device, cap=4096):
        self.count = torch.empty(int(num), dtype=torch.int32, device=device)
        self.count_max = torch.zeros(1, dtype=torch.int32, device=device)
        self.cap = int(cap)

    def c_struct(self, num, dev, wave):
        if self.count.shape[0] != num or self.count.device != dev:
            raise ValueError("FanAux was made for another bundle")
        a = _lib.pxf_program_aux()
        a.wave = wave.data_ptr() if wave is not None else None
        a.count = self.count.data_ptr()
        a.count_max = self.count_max.data_ptr()
        a.cap = self.cap
        return a

    def gratings(self):
        """Largest grating index of the launch (one 4-byte read-back); raises if some ray met no grating."""
        k = int(self.count_max.item())
        if k > self.cap:
            raise RuntimeError("gratfan: rays left after %d gratings (they never meet the array)" % self.cap)
        return k


class SegmentedProgram:
    """One launch for a bundle made of segments that run the same routines with different scalars -- a nested
    mirror assembly, one shell per segment (the reference loops over shells in Python,
    examples/axro/axialHeights.py:215-322, SMARTX.py:163-259)::

        progs = [Program().transform(0, 0, z0[k], 0, 0, 0).wolterprimary(r0[k], z0[k], 1.).reflect() ... for k]
        seg = SegmentedProgram(progs, sizes)      # sizes[k] = rays of shell k, concatenated in this order
        seg.run(bundle)

    The per-segment op tables are folded once on the host and uploaded once; ``run`` is a single kernel."""

    def __init__(self, programs, sizes, device=None):
        import numpy as np
        if len(programs) != len(sizes) or not programs:
            raise ValueError("one program per segment")
        nops = len(programs[0])
        if nops < 1 or nops > MAX_OPS:
            raise ValueError("segmented programs carry 1..%d ops" % MAX_OPS)
        for p in programs:
            if [c for c, _ in p.ops] != [c for c, _ in programs[0].ops]:
                raise ValueError("every segment must run the same routine sequence")
        self.nseg, self.nops = len(programs), nops
        self.has_vignette = programs[0].has_vignette()
        start = np.zeros(self.nseg + 1, dtype=np.int64)
        start[1:] = np.cumsum(np.asarray(sizes, dtype=np.int64))
        self.seg_start = start
        self.num = int(start[-1])
        arr = (_lib.pxf_op * (self.nseg * nops))()
        for sgm, prog in enumerate(programs):
            for k, (code, p) in enumerate(prog.ops):
                o = arr[sgm * nops + k]
                o.code = code
                for j, v in enumerate(p):
                    o.p[j] = v
        L = _lib.lib()
        nbytes = int(L.pxf_segmented_table_bytes(nops, self.nseg))
        self.table_host = torch.empty(nbytes, dtype=torch.uint8)
        _lib.check(L.pxf_segmented_table_fill(arr, nops, self.nseg, start.ctypes.data, self.table_host.data_ptr()))
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.table_dev = self.table_host.to(dev)
        self.seg_start_dev = torch.from_numpy(start).to(dev)

    def run(self, rays, alive=None, out=None):
        dev = rays[1].device
        num = rays[1].shape[0]
        if num != self.num:
            raise ValueError("bundle has %d rays, the segments add up to %d" % (num, self.num))
        if dev != self.table_dev.device:
            raise ValueError("segment table lives on %s, rays on %s" % (self.table_dev.device, dev))
        for r in rays:
            if r is not None and (not r.is_cuda or r.dtype != torch.float64 or not r.is_contiguous()
                                  or r.shape[0] != num):
                raise ValueError("ray rows must be contiguous 1-D float64 CUDA tensors of equal length")
        if self.has_vignette and alive is None:
            alive = torch.empty(num, dtype=torch.uint8, device=dev)
        ptrs = (ctypes.c_void_p * 10)(*[(r.data_ptr() if r is not None else None) for r in rays])
        optrs = None
        if out is not None:
            optrs = (ctypes.c_void_p * 10)(*[(r.data_ptr() if r is not None else None) for r in out])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pxf_trace_program_segmented(
                ptrs, optrs, num, self.table_host.data_ptr(), self.table_dev.data_ptr(),
                alive.data_ptr() if alive is not None else None, stream_ptr(dev)))
        return alive


# ------------------------------------------------------------------ transparent recording
_active = {}     # id(rays list) -> (rays, Program)


def recorder_for(rays):
    """Program currently recording for this bundle, or None."""
    ent = _active.get(id(rays))
    return ent[1] if ent is not None and ent[0] is rays else None


_alive = {}      # id(rays list) -> (rays, uint8 flags) left by the vignette predicates of the last recording


def last_alive(rays):
    """uint8 flag row written by the vignette predicates recorded for this bundle (1 = the ray passed every
    predicate), or None when the recording had none.  ``transformations.compact(rays, flags)`` removes the rest."""
    ent = _alive.get(id(rays))
    return ent[1] if ent is not None and ent[0] is rays else None


def flush(rays):
    """Execute whatever is pending for this bundle (called by anything that reads the rays)."""
    prog = recorder_for(rays)
    if prog is not None and len(prog):
        todo = Program()
        todo.ops, prog.ops = prog.ops, []
        todo._tables, prog._tables = prog._tables, {}
        alive = todo.run(rays)
        if alive is not None:
            prev = last_alive(rays)
            # (a ray stopped in an earlier launch of the same recording stays stopped)
            _alive[id(rays)] = (rays, alive if prev is None else torch.minimum(prev, alive))


class fused:
    """Context manager: calls of the ``transformations``/``surfaces`` API on ``rays`` inside the
    block are recorded and executed as one kernel on exit."""

    def __init__(self, rays):
        self.rays = rays

    def __enter__(self):
        if id(self.rays) in _active:
            raise RuntimeError("bundle is already recording")
        _active[id(self.rays)] = (self.rays, Program())
        _alive.pop(id(self.rays), None)
        return self.rays

    def __exit__(self, et, ev, tb):
        try:
            if et is None:
                flush(self.rays)
        finally:
            _active.pop(id(self.rays), None)
        return False
