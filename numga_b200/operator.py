"""Symbolic multilinear operators as SPARSE integer term lists.

An operator of arity k over subspaces (in_0 .. in_{k-1}) -> out is the set of terms
    out[o] += coeff * in_0[i_0] * ... * in_{k-1}[i_{k-1}]
stored as `index` [nnz, k+1] (columns i_0..i_{k-1}, o; rows in lexicographic order) and
`coeff` [nnz] (non-zero integers).  This is the information content of the reference's dense
`Operator.kernel` (`numga/operator/operator.py:11-21`), and `.kernel` materialises exactly that
dense array for parity checks — but composition (`fuse`, `bind`, `symmetry`, `squeeze`,
`grade_selection`) works on the term lists, so cost scales with nnz rather than n^(k+1)
(SURVEY.md §7 hard part 7).  The term list in `precompute_sparse` order
(`numga/operator/abstract.py:119-151`) is what the CUDA code generator consumes.

The semantics that decide bit-exactness of the tables are documented per method.
"""
import itertools
from functools import cached_property

import numpy as np

from numga_b200.subspace import SubSpace


def _canonical(index, coeff, ncols):
    """Merge duplicate index rows (summing coefficients), drop zeros, sort lexicographically."""
    coeff = np.asarray(coeff, dtype=np.int64).reshape(-1)
    index = np.asarray(index, dtype=np.int64).reshape(len(coeff), ncols)
    if len(coeff) == 0:
        return index, coeff
    order = np.lexsort(index.T[::-1])
    index, coeff = index[order], coeff[order]
    new_group = np.ones(len(coeff), dtype=bool)
    new_group[1:] = np.any(index[1:] != index[:-1], axis=1)
    starts = np.flatnonzero(new_group)
    coeff = np.add.reduceat(coeff, starts)
    index = index[starts]
    keep = coeff != 0
    return index[keep], coeff[keep]


_UIDS = {}          # structural key -> uid (see Operator.uid); entries are tiny and live for the process


class Operator:
    """Sparse symbolic multilinear operator; `axes` = input subspaces + output subspace."""

    def __init__(self, axes, index, coeff):
        self.axes = tuple(axes)
        self.index, self.coeff = _canonical(index, coeff, len(self.axes))

    # ---- construction helpers -------------------------------------------------------------------
    @classmethod
    def from_dense(cls, kernel, axes):
        kernel = np.asarray(kernel)
        nz = np.nonzero(kernel)
        return cls(axes, np.stack(nz, axis=1), kernel[nz])

    @classmethod
    def diagonal(cls, subspace, weights):
        n = len(subspace)
        r = np.arange(n)
        return cls((subspace, subspace), np.stack([r, r], axis=1), np.asarray(weights, dtype=np.int64))

    def _like(self, index, coeff, axes=None):
        return Operator(self.axes if axes is None else axes, index, coeff)

    # ---- basic properties ----------------------------------------------------------------------
    @cached_property
    def uid(self):
        """Small integer naming the CONTENT of this operator (axes + term table): equal operators share it, so
        compiled-kernel caches keyed on it can neither collide after garbage collection recycles an address
        (`id()`), nor miss for a structurally identical operator that was rebuilt."""
        key = (self.axes, self.index.tobytes(), self.coeff.tobytes())
        return _UIDS.setdefault(key, len(_UIDS))

    @property
    def arity(self):
        return len(self.axes) - 1

    @property
    def inputs(self):
        return self.axes[:-1]

    @property
    def output(self):
        return self.axes[-1]

    @property
    def subspace(self):
        """The 'subspace of an operator' is its output subspace."""
        return self.output

    @property
    def algebra(self):
        return self.output.algebra

    @property
    def nnz(self):
        return len(self.coeff)

    @property
    def shape(self):
        return tuple(len(a) for a in self.axes)

    @cached_property
    def kernel(self):
        """Dense integer array, identical to the reference's Operator.kernel (parity/debug only)."""
        size = int(np.prod(self.shape))
        if size > 1 << 26:
            raise MemoryError('dense kernel too large; use the term list')
        amax = int(np.abs(self.coeff).max()) if self.nnz else 0
        dense = np.zeros(self.shape, dtype=np.int8 if amax < 128 else np.int64)
        if self.nnz:
            dense[tuple(self.index.T)] = self.coeff
        return dense

    @cached_property
    def precompute_sparse(self):
        """((o, (((i0, i1, ..), coeff), ...)), ...) for each non-empty output row, terms in
        lexicographic input-index order — the layout of `operator/abstract.py:119-151`."""
        rows = {}
        for idx, c in zip(self.index.tolist(), self.coeff.tolist()):
            rows.setdefault(idx[-1], []).append((tuple(idx[:-1]), c))
        return tuple((o, tuple(rows[o])) for o in sorted(rows))

    # ---- linear structure -----------------------------------------------------------------------
    def add(self, other):
        assert self.axes == other.axes
        return self._like(np.concatenate([self.index, other.index]), np.concatenate([self.coeff, other.coeff]))

    def mul(self, s):
        return self._like(self.index, self.coeff * int(s))

    def div(self, s):
        """Integer floor division of every coefficient (operator.py:91-93)."""
        return self._like(self.index, self.coeff // int(s))

    __add__ = add
    def __sub__(self, other): return self.add(-other)
    def __neg__(self): return self.mul(-1)
    def __pos__(self): return self
    __mul__ = mul
    __rmul__ = mul

    def swapaxes(self, a, b):
        axes = list(self.axes)
        axes[a], axes[b] = axes[b], axes[a]
        index = self.index.copy()
        index[:, [a, b]] = index[:, [b, a]]
        return Operator(axes, index, self.coeff)

    # ---- composition -----------------------------------------------------------------------------
    def fuse(self, other, axis):
        """Feed the output of `self` into input `axis` of `other` (operator.py:130-138).

        The inputs of self take the place of that argument: result axes are
        other.in[:axis] + self.in + other.in[axis+1:] + (other.out,).
        """
        assert self.output is other.axes[axis], (self.output, other.axes[axis])
        # join on self.out == other.in[axis]
        by_out = {}
        for row, c in zip(self.index.tolist(), self.coeff.tolist()):
            by_out.setdefault(row[-1], []).append((row[:-1], c))
        rows, coeffs = [], []
        for orow, oc in zip(other.index.tolist(), other.coeff.tolist()):
            for srow, sc in by_out.get(orow[axis], ()):
                rows.append(orow[:axis] + srow + orow[axis + 1:])
                coeffs.append(oc * sc)
        axes = other.axes[:axis] + self.inputs + other.axes[axis + 1:]
        index = np.array(rows, dtype=np.int64).reshape(len(coeffs), len(axes))
        return Operator(axes, index, coeffs)

    def bind(self, *inputs):
        """Replace inputs of self by operators; a dict {axis: operator} binds a subset."""
        mapping = inputs[0] if isinstance(inputs[0], dict) else dict(enumerate(inputs))
        fused = self
        for axis in sorted(mapping, reverse=True):   # right-to-left keeps earlier axis numbers valid
            fused = mapping[axis].fuse(fused, axis=axis)
        return fused

    def _build(self, args):
        """Bind whichever of `args` are operators (late binding of operator-valued arguments)."""
        ops = {i: a for i, a in enumerate(args) if isinstance(a, Operator)}
        return self.bind(ops) if ops else self

    def symmetry(self, axes):
        """Declare inputs `axes` numerically identical (operator.py:108-128).

        K' = sum over permutations of those axes; if K' is divisible by the number of
        permutations everywhere, the averaged kernel is used (antisymmetric pairs cancel,
        symmetric pairs stay as two +-1 terms); otherwise the ORIGINAL kernel is kept and
        only outputs that vanish under symmetrisation are dropped.  Then squeeze.
        """
        axes = tuple(axes)
        assert len({self.axes[a] for a in axes}) == 1, 'symmetrised axes must share a subspace'
        perms = list(itertools.permutations(axes))
        idx_all, c_all = [], []
        for p in perms:
            idx = self.index.copy()
            idx[:, list(p)] = self.index[:, list(axes)]
            idx_all.append(idx)
            c_all.append(self.coeff)
        k_index, k_coeff = _canonical(np.concatenate(idx_all), np.concatenate(c_all), len(self.axes))
        if np.all(k_coeff % len(perms) == 0):
            result = self._like(k_index, k_coeff // len(perms))
        else:
            alive = np.unique(k_index[:, -1])
            keep = np.isin(self.index[:, -1], alive)
            result = self._like(self.index[keep], self.coeff[keep])
        return result.squeeze()

    def squeeze(self):
        """Drop output components that receive no term (operator.py:208-220)."""
        alive = np.unique(self.index[:, -1])
        remap = np.full(len(self.output), -1, dtype=np.int64)
        remap[alive] = np.arange(len(alive))
        index = self.index.copy()
        index[:, -1] = remap[index[:, -1]]
        return Operator(self.axes[:-1] + (self.output.slice_subspace(alive),), index, self.coeff)

    def grade_selection(self, formula):
        """Keep terms where formula(grade(in_0), .., grade(out)) holds; then squeeze (operator.py:222-231)."""
        grades = [a.grades().astype(np.int8)[self.index[:, i]] for i, a in enumerate(self.axes)]
        mask = np.broadcast_to(np.asarray(formula(*grades), dtype=bool), (self.nnz,))
        return self._like(self.index[mask], self.coeff[mask]).squeeze()

    def sum(self, axis):
        """Sum over one input axis (that input is bound to a vector of ones)."""
        axes = tuple(a for i, a in enumerate(self.axes) if i != axis)
        return Operator(axes, np.delete(self.index, axis, axis=1), self.coeff).squeeze()

    # ---- re-indexing -------------------------------------------------------------------------------
    def slice(self, inputs):
        """Restrict the leading input axes to sub-subspaces."""
        index, coeff, axes = self.index, self.coeff, list(self.axes)
        for i, sub in enumerate(inputs):
            lookup = np.full(len(axes[i]), -1, dtype=np.int64)
            lookup[axes[i].to_relative_indices(sub.blades)] = np.arange(len(sub))
            col = lookup[index[:, i]]
            keep = col >= 0
            index, coeff = index[keep].copy(), coeff[keep]
            index[:, i] = col[keep]
            axes[i] = sub
        return Operator(axes, index, coeff)

    def __getitem__(self, inputs):
        inputs = inputs if isinstance(inputs, tuple) else (inputs,)
        full = self.algebra.subspace.full()
        return self.slice(tuple(i if isinstance(i, SubSpace) else full for i in inputs))

    def select_subspace(self, output: SubSpace):
        """Re-express the output on `output`: missing blades become implicit zeros, blades not in
        `output` are dropped.  Not squeezed, on purpose (operator.py:201-203)."""
        lookup = {int(b): i for i, b in enumerate(output.blades)}
        col = np.array([lookup.get(int(self.output.blades[o]), -1) for o in self.index[:, -1]], dtype=np.int64)
        keep = col >= 0
        index = self.index[keep].copy()
        index[:, -1] = col[keep]
        return Operator(self.axes[:-1] + (output,), index, self.coeff[keep])

    def select_grade(self, k):
        return self.select_subspace(self.algebra.subspace.k_vector(k))

    # ---- golden-file text format (operator.py:246-331) ----------------------------------------------
    def cells(self):
        """{input index tuple: (sign, output blade)}; requires at most one term per input tuple."""
        out = {}
        for idx, c in zip(self.index.tolist(), self.coeff.tolist()):
            key = tuple(idx[:-1])
            assert key not in out and c in (-1, 1)
            out[key] = (c, int(self.output.blades[idx[-1]]))
        return out

    def __repr__(self):
        ins = ', '.join(a.named_str for a in self.inputs)
        return f'Operator[{self.nnz} nnz]({ins}) -> {self.output.named_str}'
