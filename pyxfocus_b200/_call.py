"""Argument marshalling shared by the f2py-shaped modules.

The reference's f2py wrappers take 1-D contiguous float64 numpy arrays and mutate the
``intent(inout)`` ones in place.  Here the arrays are rows of the
structure-of-arrays ray bundle: 1-D contiguous float64 **torch CUDA tensors** (device
pointers go straight to the C ABI, work is enqueued on torch's current stream), or --
for literal drop-in use under the reference's own Python layer -- host numpy arrays,
which are staged through the device (H2D, kernel, D2H back into the same array).
Anything else raises ``ValueError`` the way f2py does for a bad ``intent(inout)``
argument.
"""
import numpy as np
import torch

from . import _lib


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def bundle_alloc(num, device, nrows=10, zero=False):
    """Rows of one [nrows, N] fp64 allocation, every row starting on a 128-byte boundary: the double2 path needs 16,
    and a warp's 256-byte access then covers whole 32-byte sectors (rows of an odd ray count ran at 0.6 of the rate of
    aligned ones: partial-sector writes, profiles/r02_notes.md)."""
    pad = (int(num) + 15) & ~15
    base = (torch.zeros if zero else torch.empty)((nrows, max(pad, 16)), dtype=torch.float64, device=device)
    return [base[i, :num] for i in range(nrows)]


def bundle_split(rays, sizes):
    """Contiguous segments of a bundle as bundles of row VIEWS (no copies): a nested shell
    assembly keeps all shells in one allocation and traces each shell's segment with its own
    prescription (the reference loops over shells and concatenates, examples/axro/
    axialHeights.py:247-305).  Segments that start on an even ray index keep the double2 path."""
    out, lo = [], 0
    for n in sizes:
        out.append([r[lo:lo + n] for r in rays])
        lo += int(n)
    if lo != rays[1].shape[0]:
        raise ValueError("segment sizes do not add up to the bundle length")
    return out


class Staged:
    """Resolve a group of f2py-style array arguments to device pointers."""

    def __init__(self):
        self.device = None
        self.num = None
        self._writeback = []     # (numpy array, device tensor)
        self._keep = []

    def _dev(self):
        if self.device is None:
            if not torch.cuda.is_available():
                raise _lib.PxfError("pyxfocus_b200 needs a CUDA device (there is no CPU fallback)")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def _len(self, n):
        if self.num is None:
            self.num = int(n)
        elif int(n) != self.num:
            raise ValueError("shape mismatch: array length %d != num %d" % (n, self.num))

    def inout(self, a):
        """intent(inout): must already be 1-D contiguous float64; mutated in place."""
        if isinstance(a, torch.Tensor):
            if a.dtype != torch.float64 or a.dim() != 1 or not a.is_contiguous():
                raise ValueError("failed in converting argument: intent(inout) array must be a "
                                 "contiguous 1-D float64 tensor")
            if not a.is_cuda:
                raise ValueError("ray rows must live on a CUDA device (got a CPU tensor)")
            if self.device is None:
                self.device = a.device
            elif a.device != self.device:
                raise ValueError("ray rows live on different devices")
            self._len(a.shape[0])
            self._keep.append(a)
            return a.data_ptr()
        if isinstance(a, np.ndarray):
            if a.dtype != np.float64 or a.ndim != 1 or not a.flags.c_contiguous:
                raise ValueError("failed in converting argument: intent(inout) array must be a "
                                 "contiguous 1-D float64 ndarray")
            self._len(a.shape[0])
            t = torch.from_numpy(a).to(self._dev())
            self._writeback.append((a, t))
            self._keep.append(t)
            return t.data_ptr()
        raise ValueError("failed in converting argument: expected a float64 array, got %r" % type(a))

    def input(self, a, check_len=True):
        """intent(in) array: silently converted/copied like f2py does."""
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device if self.device is not None else self._dev(), dtype=torch.float64).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self._dev())
        if t.dim() != 1:
            raise ValueError("expected a rank-1 array")
        if check_len:
            self._len(t.shape[0])
        self._keep.append(t)
        return t.data_ptr()

    def mask(self, m):
        """Optional uint8 predicate (replaces the ind= gather/scatter idiom)."""
        if m is None:
            return None
        t = to_mask(m, self.num, self.device if self.device is not None else self._dev())
        self._keep.append(t)
        return t.data_ptr()

    def stream(self):
        return stream_ptr(self._dev())

    def finish(self):
        for a, t in self._writeback:
            a[...] = t.cpu().numpy()
        self._writeback = []
        self._keep = []


def to_mask(ind, num, device):
    """bool mask / index array / np.where tuple -> uint8 device mask of length num."""
    if isinstance(ind, tuple):
        if len(ind) != 1:
            raise ValueError("ind must index a 1-D array")
        ind = ind[0]
    if isinstance(ind, np.ndarray):
        ind = torch.from_numpy(np.ascontiguousarray(ind))
    elif not isinstance(ind, torch.Tensor):
        ind = torch.as_tensor(ind)
    ind = ind.to(device)
    if ind.dtype == torch.bool:
        if ind.shape[0] != num:
            raise IndexError("boolean index did not match ray count")
        return ind.contiguous().view(torch.uint8)       # a bool tensor already is one 0/1 byte per ray
    if ind.dtype == torch.uint8 and ind.shape[0] == num:
        return ind.contiguous()
    m = torch.zeros(num, dtype=torch.uint8, device=device)
    if ind.numel():
        m[ind.long()] = 1
    return m


def run(fn, st, *args):
    """Call a libpxf entry point on the staged arrays' device and stream."""
    with torch.cuda.device(st._dev()):
        rc = fn(*args)
    _lib.check(rc)
    st.finish()
