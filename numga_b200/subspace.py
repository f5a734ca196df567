"""SubSpace: the 'type' of a multivector — an ordered, immutable set of basis blades.

Mirrors the surface of the reference's `numga/subspace/{subspace,factory,namespaces}.py`
(names, argument meaning, flyweight identity, dispatch predicates).  The canonical blade
order is (grade, blade integer) — `subspace/factory.py:26-56`; that order fixes the
component layout of every multivector and is pinned bit-exactly by the golden tests.
"""
from functools import cache, cached_property

import numpy as np

from numga_b200.algebra import Algebra

_SYMMETRIC_OPS = ('reverse', 'involute', 'conjugate', 'scalar_negation', 'pseudoscalar_negation')


class _SubSpaceNamespace:
    """`subspace.inside.motor()`, `subspace.equals.scalar()`, `subspace.select.bivector()` ..."""

    def __init__(self, subspace, action):
        self._subspace = subspace
        self._factory = subspace.algebra.subspace
        self._action = action

    def __getattr__(self, name):
        target = getattr(self._factory, name)
        if isinstance(target, SubSpace):
            return self._action(self._subspace, target)
        return lambda *args: self._action(self._subspace, target(*args))

    def __getitem__(self, grades):
        grades = list(grades) if isinstance(grades, (tuple, list)) else [grades]
        return self._action(self._subspace, self._factory.from_grades(grades))

    def __call__(self, target):
        return self._action(self._subspace, target)


class SubSpace:
    """Ordered set of blades of an algebra.  Instances are flyweights: equal blade sets of one
    algebra are the same object, so identity comparison and id-hash are exact and cheap."""

    def __init__(self, algebra: Algebra, blades: np.ndarray):
        blades = np.array(blades, dtype=algebra.blade_dtype)
        assert len(np.unique(blades)) == len(blades), 'blades of a subspace must be unique'
        blades.flags.writeable = False
        self.algebra = algebra
        self.blades = blades
        self.inside = _SubSpaceNamespace(self, lambda s, t: s in t)
        self.equals = _SubSpaceNamespace(self, lambda s, t: s is t)
        self.select = _SubSpaceNamespace(self, lambda s, t: s.select_subspace(t))
        self.restrict = _SubSpaceNamespace(self, lambda s, t: s.restrict_subspace(t))

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def __len__(self):
        return len(self.blades)

    @property
    def subspace(self):
        """So a SubSpace and a symbolic Operator can be passed interchangeably."""
        return self

    @property
    def is_blade(self):
        return len(self) == 1

    def grades(self):
        return self.algebra.grade(self.blades)

    def grade(self):
        g = np.unique(self.grades())
        if len(g) != 1:
            raise Exception(f'Subspace has no unique grade: {g}')
        return int(g[0])

    # ---- printing -----------------------------------------------------------------------------
    def _bits(self, blade):
        return [(int(blade) >> k) & 1 for k in range(self.algebra.n_dimensions)]

    @cached_property
    def bit_str(self):
        return ','.join(''.join(str(b) for b in self._bits(bl)) for bl in self.blades)

    @cached_property
    def named_str(self):
        names = self.algebra.description.basis_names
        return ','.join(
            '1' if bl == 0 else ''.join(n for n, b in zip(names, self._bits(bl)) if b)
            for bl in self.blades)

    _PRETTY = ('empty', 'scalar', 'pseudoscalar', 'vector', 'bivector', 'antivector', 'antibivector',
               'trivector', 'antitrivector', 'quadvector', 'antiquadvector', 'scalar_pseudoscalar',
               'study', 'even_grade', 'odd_grade', 'reflection', 'bireflection', 'trireflection',
               'quadreflection', 'self_reverse', 'nonscalar', 'multivector')

    @cached_property
    def pretty_str(self):
        f = self.algebra.subspace
        for prefix, test in (('is_', lambda c: self is c), ('in_', lambda c: self in c)):
            for name in self._PRETTY:
                try:
                    if test(getattr(f, name)()):
                        return prefix + name
                except Exception:
                    pass
        return None

    def __repr__(self):
        return f'{self.pretty_str}: {self.named_str}'

    # ---- derived subspaces -------------------------------------------------------------------
    def _from(self, blades):
        return self.algebra.subspace.from_blades(blades)

    @cache
    def complement(self):
        return self._from(self.algebra.complement(self.blades))
    dual = complement

    def _n_zero(self):
        return self.algebra.bit_dot(self.blades, self.algebra.zeros)

    @cache
    def degenerate(self):
        return self._from(self.blades[self._n_zero() > 0])

    @cache
    def nondegenerate(self):
        return self._from(self.blades[self._n_zero() == 0])

    @cache
    def scalar_degenerate(self):
        return self.algebra.subspace.scalar().union(self.degenerate())

    @cache
    def union(self, other):
        return self._from(sorted(set(self.blades.tolist()) | set(other.blades.tolist())))

    @cache
    def intersection(self, other):
        return self._from(sorted(set(self.blades.tolist()) & set(other.blades.tolist())))

    @cache
    def difference(self, other):
        return self._from(sorted(set(self.blades.tolist()) - set(other.blades.tolist())))

    @cache
    def is_subspace(self, other):
        """True if every blade of self is in other."""
        return self is other or set(self.blades.tolist()) <= set(other.blades.tolist())

    def __contains__(self, other):
        return other.is_subspace(self)

    @cache
    def restrict_subspace(self, other):
        return self.intersection(other)

    @cache
    def select_subspace(self, other):
        return other.union(self.restrict_subspace(other))

    def slice_subspace(self, indices):
        return self._from(self.blades[indices])

    def to_relative_indices(self, blades):
        """indices such that self.blades[indices] == blades (raises when a blade is absent)."""
        blades = np.asarray(blades)
        lookup = {int(b): i for i, b in enumerate(self.blades)}
        return np.array([lookup[int(b)] for b in blades.ravel()], dtype=np.int64).reshape(blades.shape)

    # ---- subspaces of operator outputs ---------------------------------------------------------
    @cache
    def product(self, other):
        return self.algebra.operator.product(self, other).subspace

    @cache
    def wedge(self, other):
        return self.algebra.operator.wedge(self, other).subspace

    @cache
    def inner(self, other):
        return self.algebra.operator.inner(self, other).subspace

    def reject(self, other):
        return self.wedge(other).inner(other)

    @cache
    def squared(self):
        return self.algebra.operator.squared(self).subspace

    @cache
    def cubed(self):
        return self.algebra.operator.cubed(self).subspace

    @cache
    def anti_reverse_product(self, other):
        return self.algebra.operator.anti_reverse_product(self, other).subspace

    @cache
    def symmetric_alt_product(self):
        return self.algebra.operator.inverse_factor_completed_alt(self).subspace

    # ---- closure / simplicity predicates used by extension dispatch ---------------------------
    @cached_property
    def is_subalgebra(self):
        return self.product(self) in self

    @cached_property
    def minimal_subalgebra(self):
        space, prev = self, None
        while space is not prev:
            prev, space = space, space.union(self.product(space))
        return space

    @cached_property
    def minimal_exponential(self):
        space, prev = self.union(self.algebra.subspace.scalar()), None
        while space is not prev:
            prev, space = space, space.union(space.squared())
        return space

    @cache
    def is_n_simple(self, n):
        """Reduces to a scalar under n self-involution products."""
        if n == 0:
            return self.inside.scalar()
        return (self.is_squared_n_simple(n) or self.is_reverse_n_simple(n)
                or self.is_involute_n_simple(n) or self.is_conjugate_n_simple(n)
                or self.is_scalar_negation_n_simple(n) or self.is_pseudoscalar_negation_n_simple(n))

    @cache
    def is_squared_n_simple(self, n):
        return self.squared().is_n_simple(n - 1)

    @cache
    def is_alt_n_simple(self, n):
        return self.symmetric_alt_product().is_n_simple(n - 2)

    @cached_property
    def simplicity(self):
        for n in range(self.algebra.n_dimensions // 2 + 2):
            if self.is_n_simple(n):
                return n
        return None

    @cached_property
    def is_degenerate_scalar(self):
        return self is self.algebra.subspace.scalar().union(self.degenerate())

    @cached_property
    def is_degenerate(self):
        return self is self.degenerate()

    @cached_property
    def in_even(self):
        return self in self.algebra.subspace.even_grade()

    @cached_property
    def in_odd(self):
        return self in self.algebra.subspace.odd_grade()

    @cached_property
    def is_even(self):
        return self is self.algebra.subspace.even_grade()

    @cached_property
    def is_odd(self):
        return self is self.algebra.subspace.odd_grade()

    @cached_property
    def is_empty(self):
        return self is self.algebra.subspace.empty()

    @cached_property
    def in_self_reverse(self):
        return self in self.algebra.subspace.self_reverse()

    @cached_property
    def is_study(self):
        """Generalized Study number: x * scalar_negation(x) is a scalar."""
        return self.is_scalar_negation_n_simple(1)

    def __getattr__(self, name):
        """`V.wedge(B)` == `algebra.operator.wedge(V, B)` for any operator-factory name."""
        if name.startswith('__'):
            raise AttributeError(name)
        try:
            op = getattr(self.algebra.operator, name)
        except AttributeError:
            raise AttributeError(f'No operator named `{name}` defined in operator factory')
        return lambda *args: op(self, *args)


def _install_symmetric(kind):
    """x.symmetric_<kind>() / x.<kind>_product(y) / x.is_<kind>_n_simple(n), generated per involution."""
    sym_name = f'symmetric_{kind}_product'

    @cache
    def symmetric(self):
        return getattr(self.algebra.operator, sym_name)(self).subspace

    @cache
    def product(self, other):
        return getattr(self.algebra.operator, f'{kind}_product')(self, other).subspace

    @cache
    def n_simple(self, n):
        return symmetric(self).is_n_simple(n - 1)

    setattr(SubSpace, f'symmetric_{kind}', symmetric)
    setattr(SubSpace, f'{kind}_product', product)
    setattr(SubSpace, f'is_{kind}_n_simple', n_simple)


for _kind in _SYMMETRIC_OPS:
    _install_symmetric(_kind)


class SubSpaceFactory:
    """Named subspace constructors + the flyweight pool (one SubSpace object per blade set)."""

    def __init__(self, algebra: Algebra):
        self.algebra = algebra
        self.grades = np.arange(algebra.n_dimensions + 1, dtype=np.uint8)
        self._pool = {}
        all_blades = np.arange(algebra.n_blades, dtype=algebra.blade_dtype)
        g = algebra.grade(all_blades)
        self.blades_by_grade = [all_blades[g == k] for k in self.grades]

    def order_blades(self, blades):
        """Canonical order: by grade, then by blade integer."""
        blades = np.array(blades, dtype=self.algebra.blade_dtype).reshape(-1)
        return blades[np.lexsort((blades, self.algebra.grade(blades)))]

    def from_blades(self, blades) -> SubSpace:
        blades = self.order_blades(blades)
        key = blades.tobytes()
        if key not in self._pool:
            self._pool[key] = SubSpace(self.algebra, blades)
        return self._pool[key]

    def from_grades(self, grades) -> SubSpace:
        if grades is Ellipsis:
            picked = self.blades_by_grade
        else:
            grades = np.atleast_1d(np.asarray(grades)).astype(np.int64)
            picked = [self.blades_by_grade[int(g)] for g in grades]
        return self.from_blades(np.concatenate(picked) if len(picked) else [])

    def from_bit_str(self, bits: str) -> SubSpace:
        return self.from_blades([int(b[::-1], 2) for b in bits.split(',')])

    def __getattr__(self, name: str) -> SubSpace:
        """`factory.xy_zt`: subspace from '_'-separated blade names."""
        if name.startswith('_'):
            raise AttributeError(name)
        desc = self.algebra.description
        try:
            blades = [sum(1 << t for t in desc.parse_tokens(part)) for part in name.split('_')]
        except Exception:
            raise AttributeError(name)
        return self.from_blades(blades)

    @cache
    def basis(self):
        return [self.blade(b) for b in self.vector().blades]

    @cache
    def k_vector(self, k: int) -> SubSpace:
        return self.from_grades([k])

    @cache
    def blade(self, blade: int) -> SubSpace:
        return self.from_blades([blade])

    def scalar(self): return self.k_vector(0)
    def vector(self): return self.k_vector(1)
    def bivector(self): return self.k_vector(2)
    def trivector(self): return self.k_vector(3)
    def quadvector(self): return self.k_vector(4)
    def pseudoscalar(self): return self.k_vector(-1)
    def antiscalar(self): return self.k_vector(-1)
    def antivector(self): return self.k_vector(-2)
    def antibivector(self): return self.k_vector(-3)
    def antitrivector(self): return self.k_vector(-4)
    def antiquadvector(self): return self.k_vector(-5)

    def scalar_pseudoscalar(self):
        return self.from_grades([0, -1])

    def nonscalar(self):
        return self.from_grades(self.grades[1:])

    @cache
    def multivector(self) -> SubSpace:
        return self.from_grades(...)
    full = multivector

    @cache
    def empty(self) -> SubSpace:
        return self.from_blades([])

    @cache
    def even_grade(self) -> SubSpace:
        return self.from_grades(self.grades[0::2])
    motor = even_grade
    self_involute = even_grade

    @cache
    def odd_grade(self) -> SubSpace:
        return self.from_grades(self.grades[1::2])

    @cache
    def translator(self):
        return self.bireflection().difference(self.bivector().nondegenerate())

    @cache
    def rotor(self):
        return self.even_grade().nondegenerate()

    @cache
    def k_reflection(self, k: int) -> SubSpace:
        return self.scalar() if k == 0 else self.k_reflection(k - 1).product(self.vector())

    def reflection(self): return self.k_reflection(1)
    def bireflection(self): return self.k_reflection(2)
    def trireflection(self): return self.k_reflection(3)
    def quadreflection(self): return self.k_reflection(4)

    @cache
    def self_reverse(self) -> SubSpace:
        return self.from_grades(self.grades[(self.grades // 2 % 2) == 0])

    @cache
    def mod4(self) -> SubSpace:
        return self.from_grades(self.grades[0::4])

    @cache
    def degenerate(self) -> SubSpace:
        return self.full().degenerate()

    @cache
    def scalar_degenerate(self) -> SubSpace:
        return self.degenerate().union(self.scalar())

    @cache
    def nondegenerate(self) -> SubSpace:
        return self.full().difference(self.degenerate())
