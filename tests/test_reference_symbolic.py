"""The reference's symbolic-layer tests (integer logic only, no arithmetic on values), ported with their assertions:
numga/subspace/test/test_subspace.py, numga/algebra/test/test_algebra.py (2D / 3D PGA structure, sandwich and
norm-squared output grades in 8 and 9 dimensions)."""
import itertools

import numpy as np

from numga_b200.algebra import Algebra


def _grade_combinations(n_grades):
    for r in range(n_grades):
        for grades in itertools.combinations(range(n_grades), r + 1):
            yield list(grades)


def test_subspace_basic():
    """test_subspace.py:11-26: complex numbers = even subalgebra of (2,0,0)."""
    algebra = Algebra.from_pqr(2, 0, 0)
    C = algebra.subspace.even_grade()
    assert C.is_subalgebra and not C.in_self_reverse
    assert C in C and C == C
    assert C.restrict.vector().equals.empty()
    assert C.select.vector().equals.vector()
    assert isinstance(C.bit_str, str) and str(C)


def test_minimal_subalgebra():
    """test_subspace.py:29-43 (5 dimensions, every grade combination)."""
    algebra = Algebra.from_pqr(5, 0, 0)
    assert algebra.subspace.scalar().union(algebra.subspace.ab).is_subalgebra
    for grades in _grade_combinations(algebra.n_dimensions + 1):
        S = algebra.subspace.from_grades(grades)
        s = S.minimal_subalgebra
        assert s.is_subalgebra and S in s
        e = S.minimal_exponential
        assert len(e) >= 1


def test_rotor_translator_motor():
    """test_subspace.py:46-50."""
    ga = Algebra.from_pqr(3, 0, 1)
    assert ga.subspace.rotor() != ga.subspace.motor()
    assert ga.subspace.rotor().product(ga.subspace.translator()) == ga.subspace.motor()


def test_small_algebras():
    """test_algebra.py:7-51: dual numbers, complex numbers, quaternions."""
    dual = Algebra('x0')
    full = dual.subspace.full()
    c, s = dual.product(full.blades, full.blades)
    assert c.tolist() == [[0, 1], [1, 0]] and s.tolist() == [[1, 1], [1, 0]]          # x * x = 0
    cplx = Algebra('x+y+')
    even = cplx.subspace.even_grade()
    assert even.simplicity == 1
    c, s = cplx.product(even.blades, even.blades)
    assert s.tolist() == [[1, 1], [1, -1]]                                               # (xy)^2 = -1
    r3 = Algebra('x+y+z+')
    quat = r3.subspace.even_grade()
    assert quat.is_subalgebra and quat.is_n_simple(1)
    c, s = r3.product(quat.blades, quat.blades)
    assert np.all(np.diag(s) == [1, -1, -1, -1])


def test_2d_pga():
    """test_algebra.py:54-85."""
    ga = Algebra('w0x+y+')
    full, even = ga.subspace.full(), ga.subspace.even_grade()
    assert even in full and even.is_subspace(full)
    scalar, vector, bivector = ga.subspace.scalar(), ga.subspace.vector(), ga.subspace.bivector()
    rot, trans = bivector.nondegenerate(), bivector.degenerate()
    assert scalar.simplicity == 0 and vector.simplicity == 1 and bivector.simplicity == 1
    assert rot.simplicity == 1 and not rot.is_degenerate and len(rot) == 1
    assert trans.simplicity == 1 and trans.is_degenerate and len(trans) == 2
    assert even.simplicity == 1 and even.is_subalgebra
    assert ga.operator.sandwich(full, vector).output == vector


def test_3d_pga():
    """test_algebra.py:88-127."""
    ga = Algebra('x+y+z+w0')
    full, even = ga.subspace.full(), ga.subspace.even_grade()
    assert ga.subspace.k_reflection(4) == even
    assert even in full and even.is_subspace(full)
    vector, bivector, trivector = ga.subspace.vector(), ga.subspace.bivector(), ga.subspace.trivector()
    quaternion, rot, trans = even.nondegenerate(), bivector.nondegenerate(), bivector.degenerate()
    assert vector.simplicity == 1 and bivector.simplicity == 2
    assert rot.simplicity == 1 and not rot.is_degenerate and len(rot) == 3
    assert trans.simplicity == 1 and trans.is_degenerate and len(trans) == 3
    assert trivector.simplicity == 1 and even.simplicity == 2 and quaternion.simplicity == 1
    assert even.is_subalgebra


def test_sandwich_output_grades_mod_4():
    """test_algebra.py:136-156: the full sandwich of a k-vector has the input grades modulo 4 (8 dimensions)."""
    ga = Algebra('a+b+c+d+w+x+y+z+')
    grades = np.arange(ga.n_dimensions + 1)
    want = np.where(np.any(np.equal.outer(grades % 4, np.array([1]) % 4), axis=1))[0]
    S = ga.operator.full_sandwich(ga.subspace.even_grade(), ga.subspace.from_grades([1]))
    assert S.output == ga.subspace.from_grades(want)


def test_norm_squared_output_grades():
    """test_algebra.py:159-187: x ~x of a vector in 9 dimensions is a scalar."""
    ga = Algebra('a+b+c+d+e+w+x+y+z+')
    grade = 1
    antigrade = ga.n_dimensions - grade
    want = np.arange(ga.n_dimensions + 1)[::4][:min(grade, antigrade) // 2 + 1]
    kv = ga.subspace.k_vector(grade)
    S = ga.operator.reverse_product(kv, kv).symmetry((0, 1))
    assert S.output == ga.subspace.from_grades(want)
