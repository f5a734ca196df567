"""Context of the b200 backend (shape of numga/backend/torch/context.py:9-68, structure-of-arrays allocation)."""
import numpy as np
import torch

from numga.context import AbstractContext

import numga_b200
from numga_b200.backend import soa_empty, to_soa

from .multivector import B200MultiVector
from .operator import B200Operator


class B200Context(AbstractContext):
    """`mode='fast'` (FMA contraction, MUFU reciprocals; benchmarked) or `'faithful'` (bit-exact with the numpy
    backend's einsum).  `engine` is the numga_b200 context that owns kernels, caches and the tracer; numga objects
    are mapped onto it by blade set (the integer tables of both are identical, tests/test_tables_golden.py)."""

    def __init__(self, algebra, otype=B200Operator, dtype=torch.float32, device='cuda:0', mode='fast', engine=None, lazy=False):
        if dtype in (np.float32, np.float64, 'float32', 'float64'):
            dtype = {np.float32: torch.float32, np.float64: torch.float64, 'float32': torch.float32, 'float64': torch.float64}[dtype]
        super().__init__(algebra=algebra, otype=otype, dtype=dtype, mvtype=B200MultiVector)
        self.device = torch.device(device)
        self.mode = mode
        d = self.algebra.description
        self.engine = engine or numga_b200.B200Context(d.description_str, dtype=dtype, device=device, mode=mode, lazy=lazy)
        assert tuple(self.engine.algebra.description.pqr) == tuple(d.pqr)
        self._to_engine, self._from_engine = {}, {}

    # ---- mapping between numga objects and engine objects (zero copy: the tensors are shared) ------------------
    def engine_subspace(self, subspace):
        try:
            return self._to_engine[subspace]
        except KeyError:
            s = self.engine.algebra.subspace.from_blades(np.asarray(subspace.blades))
            assert np.array_equal(np.asarray(s.blades), np.asarray(subspace.blades)), 'canonical blade orders differ'
            self._to_engine[subspace] = s
            self._from_engine[s] = subspace
            return s

    def numga_subspace(self, engine_subspace):
        try:
            return self._from_engine[engine_subspace]
        except KeyError:
            s = self.algebra.subspace.from_blades(np.asarray(engine_subspace.blades))
            self._from_engine[engine_subspace] = s
            self._to_engine[s] = engine_subspace
            return s

    def to_engine(self, mv):
        if mv._emv is not None:                   # came from the engine (possibly still a recorded expression)
            return mv._emv
        mv._emv = self.engine.mvtype(self.engine, self.engine_subspace(mv.subspace), mv.values)
        return mv._emv

    def from_engine(self, emv):
        if getattr(emv, 'pending', False):        # lazy engine: nothing computed yet, keep the recorded expression
            return B200MultiVector(self, self.numga_subspace(emv.subspace), None, emv)
        return B200MultiVector(self, self.numga_subspace(emv.subspace), emv.values, emv)

    # ---- arrays (plumbing) --------------------------------------------------------------------------------------------
    def coerce_array(self, values, dtype=None):
        if isinstance(values, np.ndarray) and values.dtype == bool:
            values = values.astype(np.float64)
        t = torch.as_tensor(values, dtype=dtype or self.dtype, device=self.device)
        if t.is_floating_point() and t.dim() >= 2 and t.shape[-1] > 1 and t.stride(-1) == 1 and t[..., 0].numel() > 1:
            t = to_soa(t)              # array-of-structs input: ONE transposing copy into structure-of-arrays
        return t

    def allocate_array(self, shape):
        out = soa_empty(shape[-1], tuple(shape[:-1]), self.dtype, self.device)
        out.zero_()
        return out

    def set_array(self, arr, idx, values):
        if isinstance(idx, np.ndarray):
            idx = torch.as_tensor(idx, device=arr.device)
        arr = arr.clone()
        arr[idx] = values
        return arr

    # ---- elementwise maths on raw arrays (used by numga's generic code paths on tiny temporaries) -------------------
    def where(self, cond, a, b): return torch.where(cond, a, b)
    def nan_to_num(self, x, nan): return torch.nan_to_num(x, nan=nan)
    def abs(self, x): return torch.abs(x)
    def sqrt(self, x): return torch.sqrt(x)
    def log(self, x): return torch.log(x)
    def exp(self, x): return torch.exp(x)
    def cos(self, x): return torch.cos(x)
    def sin(self, x): return torch.sin(x)
    def pow(self, x1, x2): return torch.pow(x1, x2)
    def min(self, x, axis): return torch.min(x, dim=axis)
    def argmin(self, x, axis): return torch.argmin(x, dim=axis)
    def all(self, x, axis): return torch.all(x, dim=axis)
    def isnan(self, x): return torch.isnan(x)
    def logical_xor(self, x, y): return torch.logical_xor(x, y)
    def logical_or(self, x, y): return torch.logical_or(x, y)

    def capture(self, fn, *args, warmup=1):
        """Record every kernel `fn(*args)` launches into one CUDA graph and return it (`graph.replay()`, `graph.result`):
        numga_b200 `B200Context.capture`.  `fn` works on this backend's multivectors; with `lazy=True` it must ask for
        the `values` of what it returns, so that the recorded expressions are launched inside the capture."""
        return self.engine.capture(fn, *args, warmup=warmup)

    def realize(self, *mvs):
        """Compute the pending expressions of several multivectors with ONE kernel (lazy engine); no-op otherwise."""
        if self.engine.lazy:
            self.engine.realize(*[self.to_engine(m) for m in mvs])
        return mvs

    # ---- fusion of user code: the whole function becomes ONE kernel -------------------------------------------------
    def jit(self, fn):
        """`fn(*multivectors_or_python_scalars) -> multivector(s)`, written against the numga multivector API, compiled
        into one kernel per argument signature (numga_b200 `B200Context.jit`)."""
        inner = self.engine.jit(fn)

        def compiled(*args):
            res = inner(*[self.to_engine(a) if isinstance(a, B200MultiVector) else a for a in args])
            return tuple(self.from_engine(r) for r in res) if isinstance(res, tuple) else self.from_engine(res)
        return compiled
