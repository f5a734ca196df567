"""Operator of the b200 backend: a numga `AbstractConcreteOperator` executed by generated sm_100a kernels.

* integer kernels (the symbolic operators of numga/operator/factory.py): the term table `precompute_sparse`
  (numga/operator/abstract.py:119-151) is handed to `ngb200_operator_create`; one specialised kernel per operator and
  broadcast pattern (replaces NumpyEinsumOperator / NumpySparseOperator.__call__, backend/numpy/operator.py:117-160);
* float kernels with batch axes (results of `partial`, `inverse`, slicing: backend/numpy/operator.py:56-100): applied
  by a generated kernel that visits only the structurally non-zero entries.
"""
import numpy as np
import torch

from numga.multivector.helper import IndexHelper
from numga.operator.abstract import AbstractConcreteOperator
from numga.operator.operator import Operator
from numga.subspace.subspace import SubSpace

from numga_b200 import backend as E
from numga_b200.operator import Operator as EngineOperator


class _PendingKernel:
    """Stands in for the float kernel of a bound operator that a lazy engine has only recorded (row gathers
    `inertia_inv[idx]`, wide maps that are re-evaluated by their consumers): the shape is known, anything else computes
    the kernel."""

    def __init__(self, eop):
        self._eop = eop

    @property
    def shape(self):
        return tuple(self._eop.shape) + tuple(self._eop.kshape)

    def __getattr__(self, name):
        return getattr(self._eop.kernel, name)

    def __getitem__(self, idx):
        return self._eop.kernel[idx]


class B200Operator(AbstractConcreteOperator):

    # ---- engine counterparts ------------------------------------------------------------------------------------------
    @property
    def is_dense(self):
        return isinstance(self.operator.kernel, (torch.Tensor, _PendingKernel))

    def _engine_axes(self):
        return tuple(self.context.engine_subspace(a) for a in self.operator.axes)

    def engine_operator(self):
        """The numga_b200 operator with the same table / the same float kernel (cached on this object)."""
        op = self.__dict__.get('_engine_op')
        if op is None:
            ctx = self.context
            if self.is_dense:
                k = self.operator.kernel
                flat = k.flatten(-len(self.operator.axes))
                mask = self.__dict__.get('_mask')
                op = E.B200DenseOperator(ctx.engine, self._engine_axes(), flat, mask)
            else:
                sym = EngineOperator.from_dense(np.asarray(self.operator.kernel), self._engine_axes())
                op = E.B200Operator(ctx.engine, sym)
            self.__dict__['_engine_op'] = op
        return op

    def _wrap_dense(self, eop):
        """numga operator around an engine dense operator (float kernel [*batch, n_in.., n_out])."""
        ctx = self.context
        axes = tuple(ctx.numga_subspace(a) for a in eop.axes)
        out = type(self)(ctx, Operator(_PendingKernel(eop) if eop._flat is None else eop.kernel, axes))
        out.__dict__['_engine_op'] = eop
        out.__dict__['_mask'] = eop.mask
        return out

    # ---- numga operator interface ----------------------------------------------------------------------------------------
    @property
    def shape(self):
        """Batch shape of the kernel (backend/torch/operator.py:15-18)."""
        return tuple(self.kernel.shape)[:-(self.arity + 1)]

    def __call__(self, *inputs):
        if any(isinstance(i, SubSpace) for i in inputs):          # backend/numpy/operator.py:118-121
            return self.partial({k: i for k, i in enumerate(inputs) if not isinstance(i, SubSpace)})
        ctx = self.context
        res = self.engine_operator()(*[ctx.to_engine(i) for i in inputs])
        return ctx.from_engine(res)

    def partial(self, inputs):
        ctx = self.context
        return self._wrap_dense(self.engine_operator().partial({k: ctx.to_engine(m) for k, m in inputs.items()}))

    def inverse(self):
        assert self.arity == 1
        return self._wrap_dense(self.engine_operator().inverse())

    def transpose(self):
        assert self.arity == 1
        return self._wrap_dense(self.engine_operator().transpose())

    def sum(self, axis):
        return self._wrap_dense(self.engine_operator().sum(axis))

    def __getitem__(self, idx):
        return self._wrap_dense(self.engine_operator()[idx])

    def concatenate(self, other, axis):
        assert self.operator.axes == other.operator.axes
        return self._wrap_dense(self.engine_operator().concatenate(other.engine_operator(), axis))

    @property
    def at(self):
        def rebuild(kernel):
            return type(self)(self.context, Operator(kernel, self.operator.axes))
        return IndexHelper(rebuild, self.kernel, self.context)
