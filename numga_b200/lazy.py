"""Automatic expression fusion behind the unchanged multivector API.

With `B200Context(..., lazy=True)` every computing method of a multivector (`*`, `^`, `|`, `&`, `>>`, `~`, `+`, `-`, `/`,
`.exp()`, `.normalized()`, ...) returns a `LazyMultiVector` that only RECORDS the operation.  The recorded expression
DAG is turned into ONE generated kernel when the data is needed: on access to `.values` (indexing, `numpy()`,
reductions, passing the object to a non-lazy consumer ...), through `context.realize(a, b, ...)` (several results in
one kernel), or when the pending DAG exceeds `context.lazy_limit` operations.  This is what `context.jit(fn)` does for
an explicitly wrapped function, applied to ordinary eager-looking code: the reference's rigid-body step written
exactly like numga/examples/physics/core.py (174 array passes there) runs as a handful of kernels.

Kernels are cached on the STRUCTURE of the DAG (operation names, static arguments, argument subspaces and broadcast
pattern), so a loop that builds the same expression again re-uses the compiled kernel; python floats are runtime
parameters of the kernel, python ints are compile-time constants.
"""
import sys

import numpy as np

from numga_b200.multivector import MultiVector
from numga_b200.subspace import SubSpace

_SCALARS = (int, float, np.integer, np.floating)


class _Node:
    __slots__ = ('key', 'fn', 'args', 'static', 'count')

    def __init__(self, key, fn, args, static, count):
        self.key, self.fn, self.args, self.static, self.count = key, fn, args, static, count


def _counts(objects):
    return [(o, sys.getrefcount(o)) for o in list(objects.values())]


def _refcount_overhead():
    """References `_counts` itself holds on every object it inspects (measured, not assumed: CPython detail)."""
    probe = _Node(None, None, None, None, 0)
    return _counts({0: probe})[0][1] - 1          # minus the `probe` variable


_OVERHEAD = _refcount_overhead()


def _outside_references(inner, dag_refs):
    """[(multivector, number of references that are neither the DAG's nor this module's)]"""
    return [(mv, n - _OVERHEAD - dag_refs[id(mv)]) for mv, n in _counts(inner)]


def _method_caller(name):
    return lambda s, *a: getattr(s, name)(*a)


def make_lazy_class(base):
    """LazyMultiVector derives from the backend's concrete multivector class (so isinstance checks and every
    data-movement method keep working); `values` materialises on demand."""

    class LazyMultiVector(base):
        def __init__(self, context, subspace, node, bshape):
            self.context, self.subspace = context, subspace
            self._node, self._values, self._bshape = node, None, tuple(bshape)

        @property
        def values(self):
            if self._values is None:
                self.context.lazy_engine.realize([self])
            return self._values

        @values.setter
        def values(self, v):
            self._values, self._node = v, None

        @property
        def pending(self):
            return self._values is None

        @property
        def shape(self):
            return self._bshape if self._values is None else tuple(self._values.shape[:-1])

        def __repr__(self):
            if self._values is None:
                return f'{self.subspace.pretty_str}: <lazy, {self._node.count} pending operations>'
            return super().__repr__()

    return LazyMultiVector


class LazyEngine:
    def __init__(self, context, mv_class):
        self.context = context
        self.cls = make_lazy_class(mv_class)
        self.signatures = {}          # (name, argument descriptors) -> output subspace | None (not lazily representable)

    # ---- recording ---------------------------------------------------------------------------------------------------
    def _describe(self, a, static):
        if static or isinstance(a, SubSpace) or a is None:
            return ('static', a)
        if isinstance(a, MultiVector):
            return ('mv', a.subspace)
        if isinstance(a, (bool, int, np.integer)):
            return ('const', int(a))
        if isinstance(a, (float, np.floating)):
            return ('param',)
        return None                    # tensors / arrays of per-item factors, bound operators ...: not recorded

    def _output_subspace(self, key, fn, descr):
        """Output subspace of the operation for these argument kinds, found by tracing it once on symbols."""
        sig = (key, descr)
        if sig not in self.signatures:
            spec = []
            for d in descr:
                if d[0] == 'mv':
                    spec.append(('mv', d[1], 0))
                elif d[0] == 'param':
                    spec.append(('param',))
                else:
                    spec.append(('static', d[1]))
            self.signatures[sig] = None
            _, subspaces, single, _ = self.context.engine.trace(fn, tuple(spec))
            self.signatures[sig] = subspaces[0] if single else None
        return self.signatures[sig]

    def apply(self, key, fn, static_positions, args, out_subspace=None):
        """Record `fn(*args)` (a multivector method, or a bound operator applied to multivectors) under the hashable
        `key`; returns a LazyMultiVector, or NotImplemented when the call has to run eagerly (multiple results,
        array-valued factors)."""
        descr = tuple(self._describe(a, i in static_positions) for i, a in enumerate(args))
        if any(d is None for d in descr):
            return NotImplemented
        out = out_subspace if out_subspace is not None else self._output_subspace(key, fn, descr)
        if out is None:
            return NotImplemented
        shapes = [a.shape for a in args if isinstance(a, MultiVector)]
        bshape = shapes[0] if all(s == shapes[0] for s in shapes) else tuple(np.broadcast_shapes(*shapes))
        count = 1 + sum(a._node.count for a in args if isinstance(a, self.cls) and a._values is None)
        node = _Node(key, fn, args, descr, count)
        lazy = self.cls(self.context, out, node, bshape)
        if count > self.context.lazy_limit:
            self.realize([lazy])
        return lazy

    def apply_method(self, name, static_positions, args):
        return self.apply(name, _method_caller(name), static_positions, args)

    def gather(self, base, idx):
        """`base[idx]` for an integer index over a one-dimensional batch: nothing is copied; the kernel that consumes
        the result reads the rows through an index operand (NGB200_INDEXED)."""
        import torch
        from numga_b200.backend import Indexed
        if isinstance(base, self.cls) and base._values is None:
            if base._node.key == 'gather':
                base.values                         # gather of a gather: resolve the inner one with torch
            else:
                self.realize([base])
        if not isinstance(idx, torch.Tensor):
            idx = torch.as_tensor(np.asarray(idx), device=self.context.device)
        node = _Node('gather', None, (base, idx), (('gather',),), 0)
        lazy = self.cls(self.context, base.subspace, node, tuple(idx.shape))
        lazy._indexed = Indexed(base, idx)
        return lazy

    def scatter(self, rebuild, arr, idx, values):
        """`x.at[idx].set(values)` with pending `values`: the copy of x is made once and the kernel that computes the
        values writes them straight into its rows.  Returns None when this form does not apply."""
        import torch
        from numga_b200.backend import Indexed, _is_row_index
        if not (isinstance(values, self.cls) and values._values is None and values._node.key != 'gather'):
            return None
        if not (_is_row_index(idx) and isinstance(arr, torch.Tensor) and arr.dim() == 2):
            return None
        if not isinstance(idx, torch.Tensor):
            idx = torch.as_tensor(np.asarray(idx), device=self.context.device)
        if tuple(values.shape) != tuple(idx.shape):
            return None
        dest = rebuild(arr.clone())
        if not isinstance(dest, MultiVector) or dest.subspace is not values.subspace:
            return None
        self.realize([values], out=(Indexed(dest, idx),))
        return dest

    # ---- materialisation ------------------------------------------------------------------------------------------------
    def realize(self, targets, out=None):
        """Compute all pending `targets` (LazyMultiVectors) with ONE kernel.  Pending intermediates of the DAG that are
        still referenced from outside it (a python variable, another pending expression) are written by the same
        kernel, so nothing is computed twice later; intermediates only the DAG knows stay in registers."""
        targets = [t for t in targets if isinstance(t, self.cls) and t._values is None]
        for t in [t for t in targets if t._node.key == 'gather']:      # a bare gather asked for its values: torch does it
            from numga_b200.backend import as_soa
            base, idx = t._node.args
            t.values = as_soa(base.values[idx])
        targets = [t for t in targets if t._values is None]
        if not targets:
            return
        if len({t.shape for t in targets}) > 1:           # one kernel = one batch shape
            for t in targets:
                self.realize([t])
            return
        ctx = self.context
        bshape = targets[0].shape
        order, fns, index, leaves, leaf_index = [], [], {}, [], {}
        inner, dag_refs = {}, {}                # pending multivectors met as arguments, and how often the DAG refers to them

        def leaf(a):
            if id(a) not in leaf_index:
                leaf_index[id(a)] = len(leaves)
                leaves.append(a)
            return leaf_index[id(a)]

        def visit(mv):
            node = mv._node
            if id(node) in index:
                return index[id(node)]
            refs = []
            for a, d in zip(node.args, node.static):
                if d[0] == 'mv':
                    if isinstance(a, self.cls) and a._values is None and a._node.key == 'gather':
                        if a.shape != bshape:
                            a.values          # a gather over a smaller batch than the kernel's: resolved by torch,
                            refs.append(('l', leaf(a)))                # then broadcast like any other operand
                        else:
                            refs.append(('l', leaf(a._indexed)))       # gathered inside the kernel
                    elif isinstance(a, self.cls) and a._values is None:
                        inner[id(a)] = a
                        dag_refs[id(a)] = dag_refs.get(id(a), 0) + 1
                        refs.append(('n', visit(a)))
                    else:
                        refs.append(('l', leaf(a)))
                elif d[0] == 'param':
                    leaves.append(float(a))
                    refs.append(('l', len(leaves) - 1))
                elif d[0] == 'const':
                    refs.append(('c', d[1]))
                else:
                    refs.append(('s', d[1]))
            index[id(node)] = len(order)
            order.append((node.key, tuple(refs)))
            fns.append(node.fn)
            return index[id(node)]

        for t in targets:
            visit(t)
        target_ids = {id(t) for t in targets}
        # references to an intermediate beyond those of the DAG itself mean somebody will ask for it again
        extra = [] if out is not None else \
            [mv for mv, n in _outside_references(inner, dag_refs)
             if n > 0 and id(mv) not in target_ids and mv.shape == bshape]
        results = targets + extra
        outs = tuple(index[id(t._node)] for t in results)
        structure = (tuple(order), outs)

        def replay(*leaf_values):
            vals = []
            for (key, refs), fn in zip(order, fns):
                args = [vals[r[1]] if r[0] == 'n' else leaf_values[r[1]] if r[0] == 'l' else r[1] for r in refs]
                vals.append(fn(*args))
            result = tuple(vals[o] for o in outs)
            return result[0] if len(result) == 1 else result

        if out is not None:                       # results are written / scattered into caller-owned destinations
            ctx.engine.call(('lazy', structure), replay, tuple(leaves), out=out)
            return
        computed = ctx.engine.call(('lazy', structure), replay, tuple(leaves))
        computed = (computed,) if len(results) == 1 else computed
        for t, r in zip(results, computed):
            t.values = r.values
