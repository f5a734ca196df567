"""Non-linear multivector functions, dispatched on the subspace of their argument.

Algorithms, names, registration order and error behaviour follow the reference's
`numga/multivector/extension/{inverse,roots,logexp,norms}.py` (cited per function), because
parity with the numpy backend depends on running the SAME algorithm: e.g. `exp` is the
Cayley-seeded bisection with 15 squarings (logexp.py:94-100,118-126), not a trig closed form.

Every body is plain multivector arithmetic, so it runs unchanged on traced multivectors: on the
B200 backend each of these functions becomes ONE fused kernel with all intermediates in
registers (the reference issues 22 / 11 / 198 array passes for exp / normalized / motor_log,
SURVEY.md section 3.3).
"""
from numga_b200.multivector import new_dispatch

# ---- inverses (inverse.py) --------------------------------------------------------------------------------
motor_inverse = new_dispatch('motor_inverse', 'Inversion, assuming normalization')


@motor_inverse.register(lambda s: s.inside.motor())
def motor_inverse_normalized(m):
    return m.reverse()


inverse = new_dispatch('inverse', 'inv(x) with inv(x) * x = 1 = x * inv(x); normalization not assumed')


@inverse.register(lambda s: s.equals.empty())
def empty_inverse(s):
    """1/0 with float semantics (inverse.py:34-41)."""
    one = s.context.multivector.scalar()
    return one / (one * 0)


@inverse.register(lambda s: s.equals.scalar())
def scalar_inverse(s):
    return s.copy(1 / s.values)


def _register_involution_inverses(n):
    """x^-1 = inv(x) / (x inv(x)) whenever x inv(x) is (n-1)-simple; simplest options first (inverse.py:47-85)."""

    @inverse.register(lambda s: s.is_squared_n_simple(n))
    def squared_simple_inverse(s):
        return s / s.squared()

    @inverse.register(lambda s: s.is_reverse_n_simple(n))
    def reverse_simple_inverse(s):
        return s.reverse() / s.symmetric_reverse_product()

    @inverse.register(lambda s: s.is_conjugate_n_simple(n))
    def conjugate_simple_inverse(s):
        return s.conjugate() / s.symmetric_conjugate_product()

    @inverse.register(lambda s: s.is_scalar_negation_n_simple(n))
    def scalar_negation_simple_inverse(s):
        """The Study-number inverse."""
        return s.scalar_negation() / s.symmetric_scalar_negation_product()

    @inverse.register(lambda s: s.is_pseudoscalar_negation_n_simple(n))
    def pseudoscalar_negation_simple_inverse(s):
        return s.pseudoscalar_negation() / s.symmetric_pseudoscalar_negation_product()

    @inverse.register(lambda s: s.is_involute_n_simple(n))
    def involute_simple_inverse(s):
        return s.involute() / s.symmetric_involute_product()


_register_involution_inverses(1)


@inverse.register(lambda s: s.inside.bivector())
def bivector_inverse_5d(b):
    return b.bivector_product(-(b.symmetric_reverse_product().inverse()))


@inverse.register(lambda s: s.inside.trivector())
def trivector_inverse_5d(t):
    return t.trivector_product(t.squared().inverse())


_register_involution_inverses(2)
_register_involution_inverses(3)

# ---- norms (norms.py) ---------------------------------------------------------------------------------------
study_norm_squared = new_dispatch('study_norm_squared', 'Study norm squared')


@study_norm_squared.register(lambda s: s.is_study)
def _study_norm_squared(s):
    return s.context.operator.study_norm_squared(s.subspace)(s, s)


study_norm = new_dispatch('study_norm', 'Study norm')


@study_norm.register(lambda s: s.is_study)
def _study_norm(x):
    return x.study_norm_squared().sqrt()


norm_squared = new_dispatch('norm_squared', 'x ~x; not always positive, not always a scalar')


@norm_squared.register()
def reverse_default_norm_squared(x):
    return x.symmetric_reverse_product()


norm = new_dispatch('norm', 'sqrt(x ~x)')


@norm.register(lambda s: s.inside.scalar())
def reverse_scalar_norm(x):
    return x.abs()


@norm.register()
def reverse_default_norm(x):
    return x.norm_squared().square_root()


normalized = new_dispatch('normalized', 'Normalisation such that x ~x == 1')


@normalized.register(lambda s: s.symmetric_reverse().equals.empty())
def reverse_empty_normalized(x):
    raise Exception('Unable to normalize null objects')


@normalized.register(lambda s: s.symmetric_reverse().equals.scalar())
def reverse_scalar_normalized(x):
    """2d/3d even grade and other simple objects (norms.py:60-65)."""
    return x.symmetric_reverse_product().inverse_square_root() * x


@normalized.register(lambda s: s.symmetric_reverse().is_study and s.symmetric_reverse().restrict.nonscalar().is_degenerate)
def reverse_degenerate_study_normalized(x):
    """x ~x = s0 + ns with ns nilpotent (3D PGA motors: ns = the pseudoscalar part).  FAST mode uses the closed form
    (s0 + ns)^(-1/2) = s0^(-1/2) - ns s0^(-3/2) / 2  (one rsqrt instead of sqrt + two reciprocals; the same closed
    form as the reference's optional `extension/optimized.py:24-50`); FAITHFUL mode runs the reference's default
    chain square_root().inverse() (norms.py:66-72) operation by operation."""
    if x.context.faithful:
        return reverse_study_normalized(x)
    s = x.symmetric_reverse_product()
    s0, ns = s.restrict.scalar(), s.restrict.nonscalar()
    r = s0.inverse_square_root()
    return (r - ((ns * r) * r) * (r * 0.5)) * x          # ns r^3 / 2, multiplied in an order that cannot under/overflow early


@normalized.register(lambda s: s.symmetric_reverse().is_study)
def reverse_study_normalized(x):
    """4d/5d even grade, e.g. 3D PGA motors: x ~x = (s0, s4) is a Study number (norms.py:66-72).
    The order `s * x` matters."""
    return x.symmetric_reverse_product().square_root().inverse() * x


# ---- square roots (roots.py) ---------------------------------------------------------------------------------
square_root = new_dispatch('square_root', 'One square root s of x (s * s == x); nan for negative scalars')


@square_root.register(lambda s: s.inside.scalar())
def scalar_square_root(s):
    return s.sqrt()


@square_root.register(lambda s: s.is_study and s.restrict.nonscalar().is_degenerate)
def study_degenerate_square_root(s):
    """Study number with degenerate non-scalar part: (sqrt(s0), ns / (2 sqrt(s0))) (roots.py:19-26)."""
    s0, ns = s.restrict.scalar(), s.restrict.nonscalar()
    root = s0.sqrt()
    ci = (2 * root).inverse().nan_to_num(0)     # nan_to_num covers the numerically-zero case
    return root + ns * ci


@square_root.register(lambda s: s.is_study)
def study_square_root(s):
    s0, ns = s.restrict.scalar(), s.restrict.nonscalar()
    c = ((s0 + s.study_norm()) / 2).sqrt()
    ci = (c * 2).inverse().nan_to_num(0)
    return c + ns * ci


inverse_square_root = new_dispatch('inverse_square_root', 'invsqrt(x) * x * invsqrt(x) = 1')


@inverse_square_root.register(lambda s: s.equals.scalar())
def scalar_inverse_square_root(s):
    return s.copy(s.values ** (-0.5))


@inverse_square_root.register()
def default_inverse_square_root(x):
    return x.square_root().inverse()


motor_square_root = new_dispatch('motor_square_root', 'Square root of a NORMALIZED motor')


@motor_square_root.register(lambda s: s.inside.even_grade())
def _motor_square_root(m):
    return (m + 1).normalized()


motor_geometric_mean = new_dispatch('motor_geometric_mean', 'Geometric mean of two NORMALIZED motors')


@motor_geometric_mean.register(lambda l, r: l.inside.even_grade() and r.inside.even_grade())
def _motor_geometric_mean(l, r):
    return (l + r).normalized()


# ---- exponentials and logarithms (logexp.py) ----------------------------------------------------------------------
BISECTION_STEPS = 15    # keyword default of exp_bisect / motor_log_bisect (logexp.py:114,122)

exp = new_dispatch('exp', 'Exponential')


@exp.register(lambda s: s.equals.empty())
def empty_exp(e):
    return e.context.multivector.scalar()


@exp.register(lambda s: s.equals.scalar())
def scalar_exp(s):
    return s.copy(s.context.exp(s.values))


@exp.register(lambda s: s.is_degenerate)
def degenerate_exp(b):
    """All blades square to zero: exp(b) = 1 + b exactly."""
    return b + 1


@exp.register(lambda s: s.inside.bivector)
def default_exp(b):
    return b.exp_bisect()


log = new_dispatch('log', 'Logarithm; inverse of exp')


@log.register(lambda s: s.equals.empty())
def empty_log(e):
    raise NotImplementedError


@log.register(lambda s: s.equals.scalar())
def scalar_log(s):
    return s.copy(s.context.log(s.values))


motor_log = new_dispatch('motor_log', 'Logarithm of a NORMALIZED motor')


@motor_log.register(lambda s: s.inside.bireflection() and s.is_degenerate_scalar)
def degenerate_log(m):
    """Translators."""
    return m.restrict[2]


@motor_log.register(lambda m: m.inside.even_grade())
def default_motor_log(m):
    return m.motor_log_bisect()


motor_log_linear = new_dispatch('motor_log_linear', 'Inverse of exp_linear')


@motor_log_linear.register(lambda s: s.inside.even_grade())
def _motor_log_linear(m):
    return m.restrict[2]


exp_linear = new_dispatch('exp_linear', 'Linear exponential 1 + b; NOT normalized')


@exp_linear.register(lambda s: s.inside.bivector())
def _exp_linear(b):
    return b + 1


motor_log_linear_normalized = new_dispatch('motor_log_linear_normalized', 'Log for motors close to 1; inverse of exp_linear_normalized')


@motor_log_linear_normalized.register(lambda s: s.inside.quadreflection())
def _motor_log_linear_normalized(m):
    """'Denormalize' to scalar 1 / quadvector 0, keep the bivector part (logexp.py:79-86)."""
    return m.bivector_product(m.restrict.self_reverse().inverse())


exp_linear_normalized = new_dispatch('exp_linear_normalized', 'Exponential for small bivectors')


@exp_linear_normalized.register(lambda s: s.inside.bivector())
def _exp_linear_normalized(b):
    return (b + 1).normalized()


exp_quadratic = new_dispatch('exp_quadratic', 'Cayley map; inverse of motor_log_quadratic')


@exp_quadratic.register()
def exp_cayley(b):
    r = 1 + b / 2
    return r.squared() / r.symmetric_reverse_product()


motor_log_quadratic = new_dispatch('motor_log_quadratic', 'Inverse of exp_quadratic')


@motor_log_quadratic.register()
def motor_log_cayley(m):
    return _motor_log_linear_normalized(m.motor_square_root()) * 2


motor_log_bisect = new_dispatch('motor_log_bisect', 'Bisection logarithm; inverse of exp_bisect')


@motor_log_bisect.register(lambda s: s.inside.even_grade())
def _motor_log_bisect(m, n=BISECTION_STEPS):
    for _ in range(n):
        m = m.motor_square_root()
    return motor_log_cayley(m) * (2 ** n)


exp_bisect = new_dispatch('exp_bisect', 'Bisection exponential; inverse of motor_log_bisect')


@exp_bisect.register()
def _exp_bisect(b, n=BISECTION_STEPS):
    m = exp_cayley(b / (2 ** n))
    for _ in range(n):
        m = m.squared()
    return m


# ---- decompositions (decompose.py) ----------------------------------------------------------------------------------
decompose_polar = new_dispatch('decompose_polar', 'b = L * s: polar decomposition into a bivector line and a Study number (PGA4CS eq 68-70)')


@decompose_polar.register(lambda s: s.inside.bivector() and s.squared().inside.study())
def _decompose_polar(b):
    s = b.symmetric_reverse_product().square_root()
    si = s.inverse().nan_to_num(0)
    return b.bivector_product(si), s


decompose_invariant = new_dispatch('decompose_invariant', 'b = l + r with l.inner(r) = 0 and l*l, r*r scalar')


@decompose_invariant.register(lambda s: s.inside.bivector() and s.squared().inside.scalar())
def bivector_decompose_simple(b):
    return b, b.context.multivector.empty()


@decompose_invariant.register(lambda s: s.inside.bivector() and s.squared().is_study)
def bivector_decompose_bisimple(b):
    """decompose.py:49-53."""
    b2 = b.squared()
    s = b2.study_conjugate() / (b2.study_norm() * 2)
    return (1 / 2 + s).bivector_product(b), (1 / 2 - s).bivector_product(b)


motor_translator = new_dispatch('motor_translator', 'Translator part of a motor')


@motor_translator.register(lambda m: m.inside.rotor())
def rotor_translator(r):
    """A pure rotor has no translator part: the identity."""
    return r.context.multivector.scalar()


@motor_translator.register(lambda m: m.inside.motor())
def _motor_translator(m):
    op = m.context.operator.euclidian_factorization(m.subspace)
    return 1 + op(m, m)


motor_rotor = new_dispatch('motor_rotor', 'Rotor part of a motor')


@motor_rotor.register(lambda m: m.inside.rotor())
def rotor_rotor(r):
    return r


@motor_rotor.register(lambda m: m.inside.motor())
def _motor_rotor(m):
    return m.nondegenerate()


motor_split = new_dispatch('motor_split', 'Split a motor w.r.t. an origin o into (t, r): r leaves o invariant, t is the shortest path')


@motor_split.register(lambda m, o: m.inside.motor() and o.inside.antivector() and o.dual().is_degenerate)
def motor_split_euclidian(m, o):
    """o is the euclidean origin: m = t * r with the closed forms above (decompose.py:84-90)."""
    return m.motor_translator(), m.motor_rotor()


@motor_split.register(lambda m, o: m.inside.motor() and o.inside.antivector() and o.is_blade)
def motor_split_blade(m, o):
    """o is a single blade: drop the axes that cannot occur (decompose.py:91-99)."""
    R = m.subspace.reject(o.subspace.dual())
    T = m.context.subspace.vector().product(o.subspace.dual())
    t, r = _motor_split(m, o)
    return t.select_subspace(T), r.select_subspace(R)


@motor_split.register(lambda m, o: m.inside.motor() and o.inside.antivector())
def _motor_split(m, o):
    t = ((m >> o) / o).motor_square_root()        # shortest trajectory from o to m >> o
    return t, ~t * m                               # the residual is a rotation that leaves o invariant
