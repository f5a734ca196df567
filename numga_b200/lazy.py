"""Automatic expression fusion behind the unchanged multivector API.

With `B200Context(..., lazy=True)` every computing method of a multivector (`*`, `^`, `|`, `&`, `>>`, `~`, `+`, `-`, `/`,
`.exp()`, `.normalized()`, ...) returns a `LazyMultiVector` that only RECORDS the operation.  The recorded expression
DAG is turned into ONE generated kernel when the data is needed: on access to `.values` (indexing, `numpy()`,
reductions, passing the object to a non-lazy consumer ...), through `context.realize(a, b, ...)` (several results in
one kernel), or when the pending DAG exceeds `context.lazy_limit` operations.  This is what `context.jit(fn)` does for
an explicitly wrapped function, applied to ordinary eager-looking code: the reference's rigid-body step written
exactly like numga/examples/physics/core.py (174 array passes there) runs as SEVEN fused kernels per sub-step plus the
four row copies its `.at[].set` calls imply (measured: 20.3 ms eagerly -> 2.7 ms at 4.2 M bodies, 1.59 ms when the step is
captured into a CUDA graph with `context.capture`; the hand-arranged FusedStep takes 1.07 ms, ChainStep 0.56 ms).

Gathers `x[idx]` (integer index arrays) and scatters `x.at[idx].set(y)` become index operands of the consuming /
producing kernel; `x[j]` and `x.sum(axis=0)` over a small leading batch axis stay symbolic; a batch `[k, c]` whose
slices share sub-expressions is realised as one kernel over `c` items computing all `k` slices.

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

        def __getitem__(self, idx):
            # x[j] on the leading batch axis of a PENDING value: operations are elementwise over the batch, so the
            # slice of the result is the same expression on the slices of its leaves -- still nothing computed
            if self._values is None and isinstance(idx, (int, np.integer)) and len(self._bshape) >= 2 \
                    and self._node.key != 'gather':
                return self.context.lazy_engine.slice0(self, int(idx) % self._bshape[0])
            return super().__getitem__(idx)

        def sum(self, axis):
            n = len(self._bshape)
            if self._values is None and n >= 2 and axis in (0, -n) and self._bshape[0] <= 8 and self._node.key != 'gather':
                total = self[0]
                for j in range(1, self._bshape[0]):
                    total = total + self[j]
                return total
            return super().sum(axis)

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
            # `count` adds up the operands' counts and so over-counts shared sub-expressions (RK4 stages ...): look at
            # the actual DAG before flushing
            node.count = self._dag_size(node)
            if node.count > self.context.lazy_limit:
                self.realize([lazy])
        return lazy

    def _dag_size(self, node):
        seen, stack = set(), [node]
        while stack:
            n = stack.pop()
            if id(n) in seen or n.key == 'gather':
                continue
            seen.add(id(n))
            stack.extend(a._node for a in n.args if isinstance(a, self.cls) and a._values is None)
        return len(seen)

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

    def slice0(self, mv, j, memo=None):
        """`mv[j]` along the leading batch axis of a pending value, by rebuilding its DAG on the sliced leaves."""
        memo = {} if memo is None else memo
        key = id(mv._node)
        if key in memo:
            return memo[key]
        rank = len(mv.shape)
        node = mv._node
        if node.key == 'gather':
            base, idx = node.args
            out = self.gather(base, idx[j])
        else:
            args = []
            for a, d in zip(node.args, node.static):
                if d[0] != 'mv' or len(a.shape) < rank:
                    args.append(a)                                   # no leading axis: broadcast along it
                elif a.shape[0] == 1:
                    args.append(a[0])
                elif isinstance(a, self.cls) and a._values is None:
                    args.append(self.slice0(a, j, memo))
                else:
                    args.append(self._concrete_slice(a, j))
            count = 1 + sum(a._node.count for a in args if isinstance(a, self.cls) and a._values is None)
            out = self.cls(self.context, mv.subspace, _Node(node.key, node.fn, tuple(args), node.static, count), mv.shape[1:])
        memo[key] = out
        return out

    def _couples_leading_axis(self, targets, rank):
        """True when the pending DAG of `targets` contains values WITHOUT the leading batch axis (results of x[j],
        x.sum(axis=0) on that axis ...): the slices along that axis then share work."""
        seen, stack = set(), [t._node for t in targets]
        while stack:
            n = stack.pop()
            if id(n) in seen or n.key == 'gather':
                continue
            seen.add(id(n))
            for a in n.args:
                if isinstance(a, self.cls) and a._values is None:
                    if len(a.shape) < rank and a._node.key != 'gather':
                        return True
                    stack.append(a._node)
        return False

    def _realize_unrolled(self, targets, out):
        """A batch [k, *rest] with small k whose slices share sub-expressions (a constraint acting on its two bodies):
        ONE kernel over the [*rest] items computes all k slices, the shared part once per item -- the arrangement a
        hand-fused kernel would use -- instead of k * prod(rest) items that each recompute it."""
        from numga_b200.backend import Indexed, soa_empty
        ctx = self.context
        k = targets[0].shape[0]
        memos = [dict() for _ in range(k)]
        pieces, dests = [], []
        for n, t in enumerate(targets):
            if out is not None:
                dest = out[n]
                assert isinstance(dest, Indexed), 'unrolled scatter needs an indexed destination'
                dests.extend(Indexed(dest.target, dest.index[j]) for j in range(k))
                holder = None
            else:
                holder = ctx.mvtype(ctx, t.subspace, soa_empty(len(t.subspace), t.shape, ctx.dtype, ctx.device))
                dests.extend(ctx.mvtype(ctx, t.subspace, holder.values[j]) for j in range(k))
            pieces.extend(self.slice0(t, j, memos[j]) for j in range(k))
            t._unrolled_holder = holder
        self.realize(pieces, out=tuple(dests))
        if out is None:
            for t in targets:
                t.values = t._unrolled_holder.values
                del t._unrolled_holder

    def _expanded(self, gathered, bshape):
        from numga_b200.backend import Indexed
        cache = gathered.__dict__.setdefault('_expanded_index', {})
        if bshape not in cache:
            base, idx = gathered._node.args
            cache[bshape] = Indexed(base, idx.expand(bshape).contiguous())
        return cache[bshape]

    def _concrete_slice(self, a, j):
        """Row j of the leading batch axis of a concrete multivector: a view, no copy."""
        return self.context.mvtype(self.context, a.subspace, a.values[j])

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
        from numga_b200.backend import _identity, kernel_space
        ctx = self.context
        # the one copy `.at[].set` implies, made by a generated copy kernel straight into structure-of-arrays layout
        # (6.2 TB/s; torch's strided copy of the same array runs at 3.7 TB/s)
        copied = ctx.engine.call(('copy_rows',), _identity, (ctx.mvtype(ctx, kernel_space((arr.shape[1],)), arr),))
        dest = rebuild(copied.values)
        if not isinstance(dest, MultiVector) or dest.subspace is not values.subspace:
            return None
        self.realize([values], out=(Indexed(dest, idx),))
        return dest

    # ---- materialisation ------------------------------------------------------------------------------------------------
    def realize(self, targets, out=None):
        """Compute all pending `targets` (LazyMultiVectors) with ONE kernel.  Pending intermediates of the DAG that are
        still referenced from outside it (a python variable, another pending expression) are written by the same
        kernel, so nothing is computed twice later; intermediates only the DAG knows stay in registers.  Exception:
        values marked `_remat` (bound-operator kernels much wider than their sources, `B200Operator.partial`) are never
        written -- re-evaluating them in every consumer moves fewer bytes than reading a table would."""
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
        if len(bshape) >= 2 and bshape[0] <= 4 and self._couples_leading_axis(targets, len(bshape)):
            return self._realize_unrolled(targets, out)
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
                        if a.shape == bshape:
                            refs.append(('l', leaf(a._indexed)))       # gathered inside the kernel
                        elif len(a.shape) < len(bshape) and bshape[len(bshape) - len(a.shape):] == a.shape:
                            # the gather covers the trailing axes of the kernel's batch ([c] inside [2, c]): same
                            # rows for every leading index -> an expanded index array, still gathered in the kernel
                            refs.append(('l', leaf(self._expanded(a, bshape))))
                        else:
                            a.values          # irregular: resolved by torch, then broadcast like any other operand
                            refs.append(('l', leaf(a)))
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
             if n > 0 and id(mv) not in target_ids and mv.shape == bshape and not getattr(mv, '_remat', False)]
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
