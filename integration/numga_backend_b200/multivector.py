"""Multivector of the b200 backend: a numga `AbstractMultiVector` whose values are a torch CUDA tensor in
structure-of-arrays layout and whose methods execute as generated kernels.

numga's `AbstractMultiVector` methods are chains of operator calls plus elementwise glue on `values`
(numga/multivector/multivector.py:64-208,247-313) and its extension methods (`exp`, `motor_log`, `normalized`,
`inverse` ...; numga/multivector/extension/*.py) are chains of those: 22 / 198 / 11 array passes.  Here
every such method is overridden ON THIS SUBCLASS to run as ONE fused kernel of the numga_b200 engine -- the role
`extension/optimized.py:24-33` plays with `register(position=0)`, but scoped to this backend: the dispatch tables of
`AbstractMultiVector` are shared by all backends, so they are left untouched and numpy / jax / torch contexts in the
same process keep their own behaviour.  Methods the engine does not know fall through to numga's generic bodies, which
then run operator by operator through `B200Operator`.
"""
import numpy as np
import torch

from numga.multivector.multivector import AbstractMultiVector

from numga_b200.backend import as_soa
from numga_b200.multivector import COMPUTE_METHODS, EXTENSION_METHODS


class B200MultiVector(AbstractMultiVector):
    """`values` is a property: with a lazy engine (`B200Context(..., lazy=True)`) a result is first only the RECORDED
    expression held by the engine multivector `_emv`; asking for `values` (or anything that needs them) computes the
    whole pending chain with one generated kernel (numga_b200/lazy.py)."""

    def __init__(self, context, subspace, values=None, engine_mv=None):
        self.context, self.subspace = context, subspace
        self._values, self._emv = values, engine_mv

    @property
    def values(self) -> torch.Tensor:
        if self._values is None:
            self._values = self._emv.values
        return self._values

    @values.setter
    def values(self, v):
        self._values, self._emv = v, None

    @classmethod
    def construct(cls, context, values, subspace):
        if values is None:
            values = subspace.blades == 0          # unit scalar: the multiplicative identity
        values = context.coerce_array(values)
        assert values.shape[-1] == len(subspace)
        return cls(context=context, values=values, subspace=subspace)

    def copy(self, values=None, subspace=None):
        return self.construct(
            values=self.values.clone() if values is None else values,
            subspace=self.subspace if subspace is None else subspace, context=self.context)

    # ---- data movement (torch = plumbing; results go back to structure-of-arrays) -----------------------------------
    def __getitem__(self, idx):
        ctx = self.context
        if ctx.engine.lazy:                      # gathers / slices stay part of the recorded expression
            return ctx.from_engine(ctx.to_engine(self)[idx])
        if isinstance(idx, np.ndarray):
            idx = torch.as_tensor(idx, device=self.values.device)
        return self.copy(values=as_soa(self.values[idx]))

    @property
    def shape(self):
        return tuple(self._emv.shape) if self._values is None else tuple(self._values.shape[:-1])

    @property
    def at(self):
        """`mv.at[idx].set(rows)` (numga/multivector/helper.py): with a lazy engine the kernel that computes `rows` writes
        them straight into the copy; otherwise numga's generic helper (`context.set_array`)."""
        if not self.context.engine.lazy:
            return AbstractMultiVector.at.fget(self)
        return _LazyAt(self)

    def __len__(self):
        return self.shape[0]

    def take_along_axis(self, idx, axis):
        return self.copy(values=as_soa(torch.take_along_dim(self.values, idx[..., None], dim=axis)))

    def concatenate(self, other, axis):
        assert self.subspace == other.subspace
        return self.copy(as_soa(torch.cat([self.values, other.values], dim=axis)))

    def sum(self, axis):
        return self.context.from_engine(self.context.to_engine(self).sum(axis))

    def mean(self, axis):
        return self.context.from_engine(self.context.to_engine(self).mean(axis))

    def rearrange(self, pattern):
        from einops import rearrange
        return self.copy(as_soa(rearrange(self.values, pattern)))

    def repeat(self, pattern, **kwargs):
        from einops import repeat
        return self.copy(as_soa(repeat(self.values, pattern, **kwargs)))

    def reshape(self, shape):
        return self.copy(as_soa(self.values.reshape(tuple(shape) + (self.values.shape[-1],))))

    def flatten(self):
        return self.reshape((-1,))

    def numpy(self):
        return self.values.detach().cpu().numpy()


class _LazyAt:
    """`mv.at[idx].set(rows)` forwarded to the engine multivector (row scatter inside the kernel that computes `rows`)."""
    __slots__ = ('mv', 'idx')

    def __init__(self, mv, idx=None):
        self.mv, self.idx = mv, idx

    def __getitem__(self, idx):
        return _LazyAt(self.mv, idx)

    def set(self, rows):
        ctx = self.mv.context
        if isinstance(rows, B200MultiVector):
            rows = ctx.to_engine(rows)
        return ctx.from_engine(ctx.to_engine(self.mv).at[self.idx].set(rows))


def _engine_method(name):
    generic = getattr(AbstractMultiVector, name, None)

    def method(self, *args):
        ctx = self.context
        eargs = []
        for a in args:
            if isinstance(a, B200MultiVector):
                a = ctx.to_engine(a)
            elif isinstance(a, AbstractMultiVector):
                raise TypeError('operands belong to different backends')
            elif hasattr(a, 'blades') and hasattr(a, 'algebra'):         # a numga SubSpace argument (select / restrict)
                a = ctx.engine_subspace(a)
            eargs.append(a)
        res = getattr(ctx.to_engine(self), name)(*eargs)
        if isinstance(res, tuple):
            return tuple(ctx.from_engine(r) for r in res)
        return ctx.from_engine(res)
    method.__name__ = name
    method.__doc__ = getattr(generic, '__doc__', None)
    return method


# every computing method of the numga multivector API that the engine implements: ONE generated kernel each
for _name in COMPUTE_METHODS + tuple(EXTENSION_METHODS):
    if _name == 'combine':
        continue                    # (engine-internal signature; `+` / `-` are overridden through their dunders)
    setattr(B200MultiVector, _name, _engine_method(_name))
