"""Random multivectors for identity tests, mirroring the reference's `numga/multivector/test/util.py` (same
constructions: reflections -> bireflections -> motors), written against the backend-neutral API plus torch values."""
import numpy as np
import torch


def random_subspace(context, subspace, shape=(), rng=None):
    rng = rng or np.random
    return context.multivector(subspace=subspace, values=rng.normal(size=tuple(shape) + (len(subspace),)))


def random_normal_vector(context, shape=(), rng=None):
    return random_subspace(context, context.subspace.vector(), shape, rng)


def random_nonnegative_normal_vector(context, shape=(), rng=None):
    """Random vector with norm >= 0 (util.py:16-23): components of negative-norm vectors are inverted."""
    r = random_normal_vector(context, shape, rng)
    if context.algebra.description.pqr[1] > 0:
        n = r.squared()
        r = r.copy(values=torch.where(n.values < 0, 1 / r.values, r.values))
    return r


def random_reflection(context, shape=(), rng=None):
    r = random_nonnegative_normal_vector(context, shape, rng)
    return r * r.squared().inverse_square_root()


def random_bireflection(context, shape=(), rng=None):
    b = random_reflection(context, shape, rng) * random_reflection(context, shape, rng)
    return b * torch.sign(b.values[..., 0])


def random_2n_reflection(context, n, shape=(), rng=None):
    if n == 0:
        return context.multivector.scalar()
    return random_2n_reflection(context, n - 1, shape, rng) * random_bireflection(context, shape, rng)


def random_motor(context, shape=(), rng=None):
    m = random_2n_reflection(context, context.algebra.n_dimensions // 2, shape, rng)
    return m * torch.sign(m.values[..., 0])


def random_non_motor(context, shape=(), scale=1e-2, n=4, rng=None):
    """Not-quite-motor, not too far from unity (util.py:62-67)."""
    m = random_subspace(context, context.subspace.even_grade(), shape, rng) * scale + 1
    for _ in range(n):
        m = m.squared()
    return m


def values_of(x):
    v = x if isinstance(x, (torch.Tensor, np.ndarray)) else (x.values if hasattr(x, 'values') else x)
    return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


def assert_close(a, b, atol=1e-9):
    d = a - b if hasattr(a, 'subspace') else b - a
    if d.context.compile_only:
        return              # build-time pre-compilation pass: kernels are generated, nothing is computed
    assert np.allclose(values_of(d), 0, atol=atol), float(np.abs(values_of(d)).max())


def motor_violations(m, rng=None):
    """(m ~m - 1, grade-violating part of the full sandwich of a random vector), util.py:70-80."""
    context = m.context
    v0 = m.symmetric_reverse_product() - 1
    V = context.subspace.vector()
    v = random_subspace(context, V, m.shape, rng)
    r = context.operator.full_sandwich(m.subspace, v.subspace)(m, v, m)
    v1 = r.select_subspace(r.subspace.difference(V))
    return v0, v1


def motor_properties(m, rng=None):
    v0, v1 = motor_violations(m, rng)
    return np.linalg.norm(values_of(v0), axis=-1), np.linalg.norm(values_of(v1), axis=-1)
