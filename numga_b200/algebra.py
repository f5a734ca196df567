"""Blade bit-logic for diagonal-metric geometric algebras (table generator, integer only).

Own implementation of the integer layer the CUDA code generator consumes.  The
*outputs* (blade XOR table, sign table, reverse/involute signs) must be bit-exact
with the reference's `numga/algebra/algebra.py:95-187` and `bitops.py:11-52`;
tests/test_tables_golden.py pins that against fixtures produced by the reference.

A basis blade is an unsigned integer whose bit k says "basis vector k is present".
"""
from functools import cached_property

import numpy as np

_SIGN_CHARS = {'+': 1, '-': -1, '0': 0}
_DEFAULT_NAMES = 'abcdefghijklmnopqrstuvwxyz'   # reference default names: numga/util.py:3


def popcount(x):
    """Number of set bits of each element of an unsigned integer array."""
    x = np.asarray(x).astype(np.uint64)
    n = np.zeros(x.shape, dtype=np.int64)
    while True:
        nz = x != 0
        if not nz.any():
            return n
        n += (x & np.uint64(1)).astype(np.int64)
        x = x >> np.uint64(1)


def blade_dtype_for(n_dimensions):
    """Smallest unsigned dtype holding n_dimensions bits (matches bitops.py:18-22)."""
    nbytes = (n_dimensions - 1) // 8 + 1
    return {1: np.uint8, 2: np.uint16, 3: np.uint32, 4: np.uint32}[nbytes]


class AlgebraDescription:
    """Names and signature of the basis vectors; mirrors numga/algebra/description.py."""

    def __init__(self, basis_names, signature):
        self.basis_names = tuple(basis_names)
        self.signature = tuple(int(s) for s in signature)
        assert len(self.basis_names) == len(self.signature)
        assert all(s in (-1, 0, 1) for s in self.signature)

    @staticmethod
    def from_str(text):
        """'x+y+z+w0' -> names (x,y,z,w), signature (+,+,+,0)."""
        names, sig, start = [], [], 0
        for pos, ch in enumerate(text):
            if ch in _SIGN_CHARS:
                names.append(text[start:pos])
                sig.append(_SIGN_CHARS[ch])
                start = pos + 1
        if start != len(text) or not names:
            raise Exception(f'Cant parse algebra description {text!r}')
        return AlgebraDescription(names, sig)

    @staticmethod
    def from_signature(sig):
        return AlgebraDescription.from_str(''.join(n + s for n, s in zip(_DEFAULT_NAMES, sig)))

    @staticmethod
    def from_pqr(p, q, r):
        return AlgebraDescription.from_signature('+' * p + '-' * q + '0' * r)

    @property
    def n_dimensions(self):
        return len(self.signature)

    @property
    def n_blades(self):
        return 1 << self.n_dimensions

    @property
    def n_grades(self):
        return self.n_dimensions + 1

    @property
    def signature_str(self):
        return ''.join({1: '+', -1: '-', 0: '0'}[s] for s in self.signature)

    @property
    def description_str(self):
        return ''.join(n + s for n, s in zip(self.basis_names, self.signature_str))

    @property
    def pqr(self):
        s = self.signature
        return (s.count(1), s.count(-1), s.count(0))

    @property
    def pqr_str(self):
        return '{}{}{}'.format(*self.pqr)

    @property
    def positives(self):
        return [int(s == 1) for s in self.signature]

    @property
    def negatives(self):
        return [int(s == -1) for s in self.signature]

    @property
    def zeros(self):
        return [int(s == 0) for s in self.signature]

    def parse_tokens(self, blade_name):
        """'xy' -> (0, 1): greedy left-to-right match of basis names."""
        out, start = [], 0
        for end in range(1, len(blade_name) + 1):
            piece = blade_name[start:end]
            if piece in self.basis_names:
                out.append(self.basis_names.index(piece))
                start = end
        if start < len(blade_name):
            raise Exception(f'Unrecognized blade name {blade_name[start:]}')
        return tuple(out)

    def __mul__(self, other):
        assert not set(self.basis_names) & set(other.basis_names)
        return AlgebraDescription.from_str(self.description_str + other.description_str)


def _mask(bits):
    return sum(int(b) << k for k, b in enumerate(bits))


class Algebra:
    """Geometric algebra with a diagonal metric; blades are integer bit patterns."""

    def __init__(self, signature):
        if isinstance(signature, Algebra):
            signature = signature.description
        if isinstance(signature, AlgebraDescription):
            self.description = signature
        elif isinstance(signature, str):
            self.description = AlgebraDescription.from_str(signature)
        elif isinstance(signature, tuple):
            self.description = AlgebraDescription.from_pqr(*signature)
        else:
            raise Exception(f'Cant instantiate algebra with {signature}')
        self.blade_dtype = blade_dtype_for(self.n_dimensions)
        self.negatives = _mask(self.description.negatives)
        self.positives = _mask(self.description.positives)
        self.zeros = _mask(self.description.zeros)

    from_str = staticmethod(lambda s: Algebra(AlgebraDescription.from_str(s)))
    from_pqr = staticmethod(lambda p, q, r: Algebra(AlgebraDescription.from_pqr(p, q, r)))

    @property
    def n_dimensions(self):
        return self.description.n_dimensions

    @property
    def n_grades(self):
        return self.description.n_grades

    @property
    def n_blades(self):
        return self.description.n_blades

    def __len__(self):
        return self.n_blades

    def __mul__(self, other):
        return Algebra(self.description * other.description)

    # ---- bit logic -------------------------------------------------------------------------
    def grade(self, blades):
        return popcount(blades).astype(np.uint8)

    def bit_dot(self, a, b):
        """Number of basis vectors shared by blades a and b."""
        return popcount(np.bitwise_and(np.asarray(a, np.uint64), np.asarray(b, np.uint64)))

    def complement(self, blades):
        return np.bitwise_xor(np.asarray(blades), self.blade_dtype(self.n_blades - 1))

    def cayley(self, a, b):
        """Blade of the product a*b and the number of transpositions needed to sort it.

        swaps = #{(i, j): bit i in a, bit j in b, i > j}: every basis vector of b has to
        hop over the higher basis vectors of a.
        """
        a = np.asarray(a)
        b = np.asarray(b)
        a64, b64 = a.astype(np.uint64), b.astype(np.uint64)
        swaps = np.zeros(np.broadcast(a64, b64).shape, dtype=np.int64)
        for j in range(self.n_dimensions):
            has_j = (b64 >> np.uint64(j)) & np.uint64(1)
            above = popcount(a64 >> np.uint64(j + 1))
            swaps = swaps + has_j.astype(np.int64) * above
        return np.bitwise_xor(a, b), swaps.astype(np.int8)

    def product(self, a, b):
        """Geometric product table between blade lists a [A] and b [B].

        Returns (blade[A, B], sign[A, B] int8 in {-1, 0, +1}).
        """
        a = np.asarray(a, dtype=self.blade_dtype)
        b = np.asarray(b, dtype=self.blade_dtype)
        shared = np.bitwise_and.outer(a, b)
        n_neg = self.bit_dot(shared, self.negatives)
        n_zero = self.bit_dot(shared, self.zeros)
        blades, swaps = self.cayley(a[:, None], b[None, :])
        parity = (n_neg + swaps) & 1
        sign = np.where(n_zero == 0, 1 - 2 * parity, 0).astype(np.int8)
        return blades.astype(self.blade_dtype), sign

    def involute(self, blades):
        return (1 - 2 * (popcount(blades) & 1)).astype(np.int8)

    def reverse(self, blades):
        return (1 - 2 * ((popcount(blades) // 2) & 1)).astype(np.int8)

    @cached_property
    def pseudo_scalar_squared(self):
        pss = self.subspace.pseudoscalar().blades
        return int(self.product(pss, pss)[1][0, 0])

    # ---- factories -------------------------------------------------------------------------
    @cached_property
    def subspace(self):
        from numga_b200.subspace import SubSpaceFactory
        return SubSpaceFactory(self)

    @cached_property
    def operator(self):
        from numga_b200.opfactory import OperatorFactory
        return OperatorFactory(self)
