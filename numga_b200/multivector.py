"""MultiVector API: operator overloads and named products forwarded to the operator cache.

Mirrors `AbstractMultiVector` (numga/multivector/multivector.py:10-411): same method names,
argument meaning, operator overloads (`*`, `^`, `&`, `|`, `+`, `-`, `/`, `~`, `>>`, `<<`),
scalar upcasting, subspace-union semantics of `+`/`-`, scalar short-cut of `/`.

Two realisations share every method body below:
  * `TracedMultiVector` — `values` is a `SymArray`; calling a method appends instructions to the
    active trace (this is how whole chains become ONE kernel);
  * `B200MultiVector` (numga_b200/backend.py) — `values` is a torch CUDA tensor in
    structure-of-arrays layout; each public method is compiled once per (method, argument
    subspaces, broadcast pattern) by tracing the body here, and then launched as one kernel.
"""
from numga_b200.dispatch import SubspaceDispatch
from numga_b200.subspace import SubSpace


class _IndexSetter:
    """`x.at[idx].set(values)` (numga/multivector/helper.py:2-23)."""

    def __init__(self, rebuild, arr, context, idx=None):
        self._rebuild, self._arr, self._context, self._idx = rebuild, arr, context, idx

    def __getitem__(self, idx):
        return _IndexSetter(self._rebuild, self._arr, self._context, idx)

    def set(self, values):
        lazy = getattr(self._context, 'lazy_engine', None)
        if lazy is not None and self._context.lazy and isinstance(values, MultiVector):
            scattered = lazy.scatter(self._rebuild, self._arr, self._idx, values)
            if scattered is not None:
                return scattered                      # rows written in place by the kernel that computes them
        if isinstance(values, MultiVector):
            values = values.values
        return self._rebuild(self._context.set_array(self._arr, self._idx, values))


class _MVNamespace:
    """`mv.select.bivector()`, `mv.select[2]`, `mv.restrict.self_reverse()` ... (multivector/namespaces.py)."""

    def __init__(self, mv, method):
        self._mv, self._method = mv, method
        self._factory = mv.context.subspace

    def __getattr__(self, name):
        target = getattr(self._factory, name)
        if isinstance(target, SubSpace):
            return getattr(self._mv, self._method)(target)
        return lambda *args: getattr(self._mv, self._method)(target(*args))

    def __getitem__(self, grades):
        grades = list(grades) if isinstance(grades, tuple) else [grades]
        return getattr(self._mv, self._method)(self._factory.from_grades(grades))


class MultiVector:
    """values[..., k] is the coefficient of blade subspace.blades[k]; leading axes are the batch."""

    def __init__(self, context, subspace, values):
        self.context = context
        self.subspace = subspace
        self.values = values

    @property
    def algebra(self):
        return self.subspace.algebra

    @property
    def operator(self):
        return self.context.operator

    def copy(self, values=None, subspace=None):
        return self.context.multivector(
            values=self.context.copy_array(self.values) if values is None else values,
            subspace=self.subspace if subspace is None else subspace)

    @property
    def at(self):
        return _IndexSetter(lambda v: self.copy(values=v), self.values, self.context)

    def __repr__(self):
        return f'{self.subspace.pretty_str}: <{self.subspace.named_str}>\n{self.values}'

    # ---- subspace selection ---------------------------------------------------------------------
    @property
    def select(self):
        """Components inside the selected subspace; absent ones are implicit zeros."""
        return _MVNamespace(self, 'select_subspace')

    @property
    def restrict(self):
        """Components present in both the selected subspace and this multivector."""
        return _MVNamespace(self, 'restrict_subspace')

    def select_subspace(self, subspace):
        return self.operator.select(self.subspace, subspace)(self)

    def restrict_subspace(self, subspace):
        return self.operator.restrict(self.subspace, subspace)(self)

    def select_grade(self, k):
        return self.select_subspace(self.algebra.subspace.k_vector(k))

    def restrict_grade(self, k):
        return self.restrict_subspace(self.algebra.subspace.k_vector(k))

    # ---- linear combination ----------------------------------------------------------------------
    def upcast(self, x):
        """Scalars (python numbers, arrays of per-item factors) become grade-0 multivectors."""
        if isinstance(x, MultiVector):
            return x
        return self.context.multivector.scalar(self.context.coerce_array(x)[..., None])

    def combine(self, other, sign):
        """self + sign * other on the union of the two subspaces (multivector.py:247-261)."""
        other = self.upcast(other)
        if self.subspace is other.subspace:
            return self.copy(values=self.values + other.values * sign)
        union = self.subspace.union(other.subspace)
        return self.context.multivector(
            values=self.select_subspace(union).values + other.select_subspace(union).values * sign,
            subspace=union)

    def __neg__(self): return self.copy(-self.values)
    def __pos__(self): return self
    def __invert__(self): return self.reverse()
    def __mul__(self, other): return self.product(self.upcast(other))
    def __rmul__(self, other): return self.upcast(other).product(self)
    def __xor__(self, other): return self.wedge(other)
    def __and__(self, other): return self.regressive(other)
    def __or__(self, other): return self.inner(other)
    def __add__(self, other): return self.combine(other, +1)
    def __radd__(self, other): return self.combine(other, +1)
    def __sub__(self, other): return self.combine(self.upcast(other), -1)
    def __rsub__(self, other): return self.upcast(other).combine(self, -1)
    def __rshift__(self, other): return self.sandwich(other)
    def __lshift__(self, other): return self.reverse_sandwich(other)
    def __truediv__(self, other): return self._divide(self.upcast(other))
    def __rtruediv__(self, other): return self.upcast(other)._divide(self)

    def _divide(self, other):
        if other.subspace.equals.scalar():
            return self.copy(values=self.values / other.values)   # no detour through 1/other
        return self * other.inverse()

    # ---- inverses that are plain formulas (multivector.py:315-343) ---------------------------------
    def inverse_hitzer(self):
        i = self.inverse_factor()
        return i / self.scalar_product(i)

    def inverse_shirokov(self):
        """Faddeev-LeVerrier style inverse that works in any dimension (arXiv:2005.04015, thm 4)."""
        n = 2 ** ((self.algebra.n_dimensions + 1) // 2)
        u = self
        for k in range(1, n):
            u_minus_c = u - (n / k) * u.select[0]
            u = self * u_minus_c
        return u_minus_c / u.restrict[0]

    # ---- partially bound operators -------------------------------------------------------------------
    def sandwich_map(self, v: SubSpace = None):
        """The sandwich self >> . as a linear map on subspace v (motor -> matrix)."""
        v = self.algebra.subspace.vector() if v is None else v
        return self.operator.sandwich(self.subspace, v).partial({0: self, 2: self})

    def inertia_map(self, b: SubSpace = None):
        """Rates -> momenta for the points in self:  P v (P x B)."""
        b = self.algebra.subspace.bivector() if b is None else b
        return self.operator.inertia(self.subspace, b).partial({0: self, 1: self})

    # ---- elementwise maths ------------------------------------------------------------------------------
    def sqrt(self): return self.copy(values=self.values ** 0.5)
    def invsqrt(self): return self.copy(values=self.values ** -0.5)
    def nan_to_num(self, nan): return self.copy(self.context.nan_to_num(self.values, nan))
    def abs(self): return self.copy(self.context.abs(self.values))
    def cos(self): return self.copy(values=self.context.cos(self.values))
    def sin(self): return self.copy(values=self.context.sin(self.values))


# ---- generated forwarding methods: mv.<name>(...) -> context.operator.<op>(subspaces...)(mvs...) ------------
_UNARY = {
    'dual': 'right_hodge', 'dual_inverse': 'right_hodge_inverse', 'reverse': 'reverse', 'involute': 'involute',
    'conjugate': 'conjugate', 'study_conjugate': 'study_conjugate', 'scalar_negation': 'scalar_negation',
    'pseudoscalar_negation': 'pseudoscalar_negation', 'degenerate': 'degenerate', 'nondegenerate': 'nondegenerate',
}
_BINARY = {
    'product': 'product', 'anti_product': 'anti_product', 'wedge': 'wedge', 'outer': 'wedge',
    'anti_wedge': 'anti_wedge', 'regressive': 'regressive', 'cross': 'cross_product', 'inner': 'inner',
    'anti_inner': 'anti_inner', 'scalar_product': 'scalar_product', 'bivector_product': 'bivector_product',
    'commutator': 'commutator', 'commutator_anti': 'commutator_anti_product',
    'anti_commutator_anti': 'anti_commutator_anti_product', 'anti_commutator': 'anti_commutator_product',
    'commutator_product': 'commutator_product', 'commutator_anti_product': 'commutator_anti_product',
    'anti_commutator_anti_product': 'anti_commutator_anti_product', 'anti_commutator_product': 'anti_commutator_product',
    'left_contract': 'left_contract', 'right_contract': 'right_contract', 'reverse_product': 'reverse_product',
    'involute_product': 'involute_product', 'conjugate_product': 'conjugate_product',
    'scalar_negation_product': 'scalar_negation_product', 'pseudoscalar_negation_product': 'pseudoscalar_negation_product',
    'anti_reverse_product': 'anti_reverse_product',
}
_QUADRATIC = ('squared', 'symmetric_reverse_product', 'symmetric_involute_product', 'symmetric_conjugate_product',
              'symmetric_scalar_negation_product', 'symmetric_pseudoscalar_negation_product')
_TERNARY_SELF = {   # op(self, other, self)
    'sandwich': 'sandwich', 'reverse_sandwich': 'reverse_sandwich', 'full_sandwich': 'full_sandwich',
    'transform': 'transform', 'reverse_transform': 'reverse_transform',
}


def _install():
    def unary(op):
        return lambda self: getattr(self.operator, op)(self.subspace)(self)

    def binary(op):
        return lambda self, other: getattr(self.operator, op)(self.subspace, other.subspace)(self, other)

    def quadratic(op):
        return lambda self: getattr(self.operator, op)(self.subspace)(self, self)

    def ternary(op):
        return lambda self, other: getattr(self.operator, op)(self.subspace, other.subspace)(self, other, self)

    for name, op in _UNARY.items():
        setattr(MultiVector, name, unary(op))
    for name, op in _BINARY.items():
        setattr(MultiVector, name, binary(op))
    for name in _QUADRATIC:
        setattr(MultiVector, name, quadratic(name))
    for name, op in _TERNARY_SELF.items():
        setattr(MultiVector, name, ternary(op))
    MultiVector.inverse_factor = lambda self: self.operator.inverse_factor(self.subspace)(self, self, self)
    MultiVector.trivector_product = lambda self, other: self.operator.n_product(
        self.subspace, other.subspace, self.context.subspace.k_vector(3))(self, other)


_install()

# names of the methods that compute (as opposed to move data); the concrete backend compiles these
COMPUTE_METHODS = (
    tuple(_UNARY) + tuple(_BINARY) + _QUADRATIC + tuple(_TERNARY_SELF) +
    ('inverse_factor', 'trivector_product', 'select_subspace', 'restrict_subspace', 'select_grade', 'restrict_grade',
     'combine', '__neg__', '__invert__', '__mul__', '__rmul__', '__xor__', '__and__', '__or__', '__add__', '__radd__',
     '__sub__', '__rsub__', '__rshift__', '__lshift__', '__truediv__', '__rtruediv__', 'inverse_hitzer',
     'inverse_shirokov', 'sqrt', 'invsqrt', 'nan_to_num', 'abs', 'cos', 'sin'))


class TracedMultiVector(MultiVector):
    """Multivector whose values are symbolic; exists only while a function is being traced."""

    @classmethod
    def construct(cls, context, values, subspace):
        assert len(values) == len(subspace), (len(values), len(subspace))
        return cls(context, subspace, values)


def new_dispatch(name, doc=''):
    """Create a subspace-dispatched extension method on MultiVector (see extension.py)."""
    d = SubspaceDispatch(doc)
    setattr(MultiVector, name, d)
    EXTENSION_METHODS.append(name)
    return d


EXTENSION_METHODS = []
