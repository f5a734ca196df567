"""Named constructors for the symbolic operators (the "which products exist" catalogue).

Same names, aliases, argument meaning and output subspaces as the reference's
`numga/operator/factory.py` (cited per method), built on the sparse term-list `Operator`.
Every constructor accepts either a SubSpace or an already-built Operator per argument;
an Operator argument is composed in ("late binding", factory.py:1-24).

The resulting integer tables are pinned bit-exactly against the reference by
tests/test_tables_golden.py.
"""
from functools import cache

import numpy as np

from numga_b200.algebra import Algebra
from numga_b200.operator import Operator
from numga_b200.subspace import SubSpace


def _parity_sign(p):
    return 1 - 2 * (np.asarray(p).astype(np.int64) & 1)


# grade-selection rules of the graded products: (grade_l, grade_r, grade_out) -> keep?
_GRADE_RULES = {
    'geometric_product': lambda l, r, o: True,                      # factory.py:233-237
    'wedge_product': lambda l, r, o: (l + r) == o,                  # factory.py:240-244
    'inner_product': lambda l, r, o: np.abs(l - r) == o,            # factory.py:245-251
    'scalar_product': lambda l, r, o: 0 == o,                       # factory.py:252-256
    'left_contraction_product': lambda l, r, o: (r - l) == o,       # factory.py:273-278
    'right_contraction_product': lambda l, r, o: (l - r) == o,      # factory.py:279-283
}

_ALIASES = {
    'geometric_product': ('gp', 'product'),
    'wedge_product': ('op', 'outer', 'wedge', 'outer_product'),
    'inner_product': ('ip', 'inner'),
    'scalar_product': ('sp', 'dp', 'dot', 'scalar', 'dot_product'),
    'left_contraction_product': ('lc', 'left_contract'),
    'right_contraction_product': ('rc', 'right_contract'),
    'left_interior_product': ('lip', 'left_interior'),
    'right_interior_product': ('rip', 'right_interior'),
    'anti_geometric_product': ('agp', 'anti_product'),
    'anti_wedge_product': ('aop', 'anti_outer', 'anti_wedge', 'anti_outer_product'),
    'anti_inner_product': ('aip', 'anti_inner'),
    'anti_scalar_product': ('asp', 'adp', 'anti_dot', 'anti_dot_product', 'anti_scalar'),
    'anti_left_interior_product': ('alip', 'anti_left_interior'),
    'anti_right_interior_product': ('arip', 'anti_right_interior'),
    'regressive_product': ('rp', 'regressive'),
    'commutator_product': ('commutator',),
    'anti_commutator_product': ('anti_commutator',),
    'commutator_anti_product': ('commutator_anti',),
    'anti_commutator_anti_product': ('anti_commutator_anti',),
    'right_hodge': ('dual',),
    'right_hodge_inverse': ('dual_inverse',),
}

_INVOLUTIONS = ('reverse', 'involute', 'conjugate', 'scalar_negation', 'pseudoscalar_negation', 'study_conjugate')


class OperatorFactory:
    """Builds and caches symbolic operators, keyed on their subspace (or operator) arguments."""

    def __init__(self, algebra: Algebra):
        self.algebra = algebra

    def build(self, operator: Operator, args):
        return operator._build(args)

    # ---- unary, diagonal ------------------------------------------------------------------------------
    def diagonal(self, v, weights) -> Operator:
        return Operator.diagonal(v.subspace, weights)._build((v,))

    @cache
    def negatives(self, v):
        """Sign each blade of v picks up when multiplied with itself through negative basis vectors."""
        return _parity_sign(self.algebra.bit_dot(v.subspace.blades, self.algebra.negatives))

    @cache
    def identity(self, v) -> Operator:
        return self.diagonal(v, np.ones(len(v.subspace), dtype=np.int64))

    @cache
    def select(self, i, o) -> Operator:
        """Re-express on subspace o: shared blades copied, missing ones zero; rows are NOT squeezed
        (factory.py:80-89) so that broadcasting to implicit zeros keeps working."""
        i, o = i.subspace, o.subspace
        pos = {int(b): k for k, b in enumerate(o.blades)}
        pairs = [(k, pos[int(b)]) for k, b in enumerate(i.blades) if int(b) in pos]
        index = np.array(pairs, dtype=np.int64).reshape(len(pairs), 2)
        return Operator((i, o), index, np.ones(len(pairs), dtype=np.int64))

    @cache
    def restrict(self, i, o) -> Operator:
        return self.select(i, o.subspace.intersection(i.subspace))

    @cache
    def nondegenerate(self, v) -> Operator:
        return self.select(v, v.subspace.nondegenerate())

    @cache
    def degenerate(self, v) -> Operator:
        return self.select(v, v.subspace.degenerate())

    @cache
    def reverse(self, v) -> Operator:
        return self.diagonal(v, self.algebra.reverse(v.subspace.blades))

    @cache
    def involute(self, v) -> Operator:
        return self.diagonal(v, self.algebra.involute(v.subspace.blades))

    @cache
    def conjugate(self, v) -> Operator:
        return self.involute(self.reverse(v))

    @cache
    def scalar_negation(self, v) -> Operator:
        return self.diagonal(v, _parity_sign(v.subspace.grades() > 0))

    study_conjugate = scalar_negation

    @cache
    def pseudoscalar_negation(self, v) -> Operator:
        """Flips every grade except 0 and n (factory.py:100-109)."""
        g = v.subspace.grades()
        keep = (g == 0) | (g == self.algebra.n_dimensions)
        return self.diagonal(v, -_parity_sign(keep))

    # ---- unary, complements and duals ---------------------------------------------------------------
    def _signed_permutation(self, v, blades, signs) -> Operator:
        """v.blades[k] -> signs[k] * blades[k]; output subspace is the complement of v."""
        sub = v.subspace
        out = sub.complement()
        index = np.stack([np.arange(len(sub)), out.to_relative_indices(blades)], axis=1)
        return Operator((sub, out), index, signs)._build((v,))

    @property
    def _pss(self):
        return self.algebra.subspace.pseudoscalar().blades

    @cache
    def right_complement(self, v) -> Operator:
        """d(v) = v * I with metric signs (factory.py:142-148)."""
        blades, swaps = self.algebra.cayley(v.subspace.blades, self._pss)
        return self._signed_permutation(v, blades, _parity_sign(swaps) * self.negatives(v.subspace))

    @cache
    def left_complement(self, v) -> Operator:
        blades, swaps = self.algebra.cayley(self._pss, v.subspace.blades)
        return self._signed_permutation(v, blades, _parity_sign(swaps) * self.negatives(v.subspace))

    @cache
    def right_complement_dual(self, v) -> Operator:
        """d(v) with v * d(v) = I, metric-free (factory.py:156-163)."""
        blades = np.bitwise_xor(self._pss, v.subspace.blades)
        _, swaps = self.algebra.cayley(v.subspace.blades, blades)
        return self._signed_permutation(v, blades, _parity_sign(swaps))

    @cache
    def left_complement_dual(self, v) -> Operator:
        blades = np.bitwise_xor(self._pss, v.subspace.blades)
        _, swaps = self.algebra.cayley(blades, v.subspace.blades)
        return self._signed_permutation(v, blades, _parity_sign(swaps))

    @cache
    def left_hodge(self, v) -> Operator:
        return self.left_complement_dual(self.diagonal(v, self.negatives(v)))

    @cache
    def right_hodge(self, v) -> Operator:
        return self.right_complement_dual(self.diagonal(v, self.negatives(v)))

    def unary_inverse(self, op: Operator) -> Operator:
        """Inverse of a signed-permutation unary operator (factory.py:185-189 uses linalg.inv)."""
        assert op.arity == 1
        kernel = np.linalg.inv(op.kernel).astype(op.kernel.dtype)
        return Operator.from_dense(kernel, op.axes[::-1])

    @cache
    def right_hodge_inverse(self, v) -> Operator:
        return self.unary_inverse(self.right_hodge(v.subspace.complement()))._build((v,))

    @cache
    def left_hodge_inverse(self, v) -> Operator:
        return self.unary_inverse(self.left_hodge(v.subspace.complement()))._build((v,))

    # ---- binary: graded products ---------------------------------------------------------------------
    def make_product(self, l, r, formula) -> Operator:
        """Geometric product restricted by a grade rule (factory.py:218-231): the rule is applied
        on the raw blade table, before any operator-valued argument is bound."""
        ls, rs = l.subspace, r.subspace
        full = self.algebra.subspace.full()
        blades, sign = self.algebra.product(ls.blades, rs.blades)
        li, ri = np.nonzero(sign)
        index = np.stack([li, ri, full.to_relative_indices(blades[li, ri])], axis=1)
        raw = Operator((ls, rs, full), index.reshape(-1, 3), sign[li, ri])
        return raw.grade_selection(formula)._build((l, r)).squeeze()

    @cache
    def n_product(self, l, r, o) -> Operator:
        """Geometric product with the output restricted to subspace o (factory.py:264-269)."""
        gp = self.geometric_product(l, r)
        return self.restrict(gp.subspace, o).bind(gp)

    @cache
    def bivector_product(self, l, r) -> Operator:
        return self.n_product(l, r, self.algebra.subspace.bivector())

    # interior / anti products (factory.py:287-346)
    @cache
    def left_interior_product(self, l, r) -> Operator:
        return self.anti_wedge_product(self.left_complement(l), r).squeeze()

    @cache
    def right_interior_product(self, l, r) -> Operator:
        return self.anti_wedge_product(l, self.right_complement(r)).squeeze()

    def antify(self, op, *inputs) -> Operator:
        """left_complement(op(right_complement(x) ...))"""
        return self.left_complement(op(*(self.right_complement(i) for i in inputs))).squeeze()

    @cache
    def regressive_product(self, l, r) -> Operator:
        """dual_inverse(dual(l) ^ dual(r)) (factory.py:348-357)."""
        return self.right_hodge_inverse(self.wedge_product(self.right_hodge(l), self.right_hodge(r))).squeeze()

    @cache
    def cross_product(self, l, r) -> Operator:
        return self.right_hodge(self.wedge_product(l, r)).squeeze()

    # commutators (factory.py:371-396); deliberately uncached like the reference
    def _commutator(self, op, sign, l, r) -> Operator:
        return (op(l, r) + (op(r, l) * sign).swapaxes(0, 1)).div(2).squeeze()

    def commutator_product(self, l, r):
        return self._commutator(self.geometric_product, -1, l, r)

    def anti_commutator_product(self, l, r):
        return self._commutator(self.geometric_product, +1, l, r)

    def commutator_anti_product(self, l, r):
        return self._commutator(self.anti_geometric_product, -1, l, r)

    def anti_commutator_anti_product(self, l, r):
        return self._commutator(self.anti_geometric_product, +1, l, r)

    @cache
    def anti_reverse_product(self, l, r) -> Operator:
        return self.antify(self.reverse_product, l, r)

    # ---- quadratic / higher forms of one argument -------------------------------------------------
    @cache
    def squared(self, v) -> Operator:
        return self.geometric_product(v, v).symmetry((0, 1))

    @cache
    def cubed(self, v) -> Operator:
        return self.geometric_product(self.squared(v), v).symmetry((0, 1, 2))

    @cache
    def study_norm_squared(self, v) -> Operator:
        return self.scalar_product(v, self.study_conjugate(v))

    @cache
    def full_sandwich(self, R, v) -> Operator:
        """(R * v) * ~R with both R slots declared identical (factory.py:495-507)."""
        return self.geometric_product(self.geometric_product(R, v), self.reverse(R)).symmetry((0, 2))

    @cache
    def sandwich(self, R, v) -> Operator:
        """Grade-preserving part of the full sandwich (factory.py:519-528); not divided by |R|^2."""
        grades = np.unique(v.subspace.grades())
        return self.full_sandwich(R, v).grade_selection(
            lambda l, m, r, o: np.any(np.asarray(o)[..., None] == grades, axis=-1))

    @cache
    def reverse_sandwich(self, R, v) -> Operator:
        return self.sandwich(self.reverse(R), v)

    @cache
    def transform(self, R, v) -> Operator:
        """Sandwich with the sign that preserves the outermorphism.  (The reference's version,
        factory.py:535-545, raises on a shape mismatch; this is the evident intent.)"""
        parity = np.unique(R.subspace.grades() % 2)
        if len(parity) != 1:
            raise Exception(f'Not all elements are equal: {parity}')
        S = self.sandwich(R, v)
        sign = _parity_sign(S.axes[1].grades().astype(np.int64) * int(parity[0]))
        return Operator(S.axes, S.index, S.coeff * sign[S.index[:, 1]])

    @cache
    def inverse_transform(self, R, v) -> Operator:
        return self.transform(self.reverse(R), v)
    reverse_transform = inverse_transform

    @cache
    def inverse_factor(self, x) -> Operator:
        p = self.symmetric_reverse_product(x)
        return self.geometric_product(self.reverse(x), self.conjugate(p)).symmetry((0, 1, 2))

    @cache
    def inverse_factor_completed(self, x) -> Operator:
        p = self.symmetric_reverse_product(x)
        return self.geometric_product(p, self.conjugate(p)).symmetry((0, 1, 2, 3))

    @cache
    def inverse_factor_alt(self, x) -> Operator:
        p = self.symmetric_reverse_product(x)
        return self.geometric_product(self.reverse(x), self.scalar_negation(p)).symmetry((0, 1, 2))

    @cache
    def inverse_factor_completed_alt(self, x) -> Operator:
        p = self.symmetric_reverse_product(x)
        return self.geometric_product(p, self.scalar_negation(p)).symmetry((0, 1, 2, 3))

    def euclidian_factorization(self, m) -> Operator:
        """Translational bivector part of a motor w.r.t. the euclidean origin: <dual(m * dual(m))>_2, both slots
        identical (factory.py:627-633); uncached like the reference."""
        op = self.geometric_product(m, self.right_hodge(m))
        return self.right_hodge(op).select_grade(2).symmetry((0, 1))

    @cache
    def inertia(self, l, r) -> Operator:
        """l v (l x r): maps rates to momenta for points l (factory.py:599-602)."""
        return self.regressive_product(l, self.commutator_product(l, r)).symmetry((0, 1))


def _install_generated():
    def graded(name, rule):
        @cache
        def product(self, l, r):
            return self.make_product(l, r, rule)
        product.__name__ = name
        return product

    for name, rule in _GRADE_RULES.items():
        setattr(OperatorFactory, name, graded(name, rule))

    def anti(base):
        @cache
        def product(self, l, r):
            return self.antify(getattr(self, base), l, r)
        return product

    for base in ('geometric_product', 'wedge_product', 'inner_product', 'scalar_product',
                 'left_interior_product', 'right_interior_product'):
        setattr(OperatorFactory, 'anti_' + base, anti(base))

    def involution_products(kind):
        @cache
        def binary(self, l, r):
            """l * kind(r)"""
            return self.geometric_product(l, getattr(self, kind)(r))

        @cache
        def symmetric(self, x):
            """x * kind(x), both slots identical"""
            return binary(self, x, x).symmetry((0, 1))
        return binary, symmetric

    for kind in _INVOLUTIONS:
        binary, symmetric = involution_products(kind)
        setattr(OperatorFactory, f'{kind}_product', binary)
        setattr(OperatorFactory, f'symmetric_{kind}_product', symmetric)

    for name, aliases in _ALIASES.items():
        for alias in aliases:
            setattr(OperatorFactory, alias, getattr(OperatorFactory, name))


_install_generated()
