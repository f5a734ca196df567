"""The drop-in backend `integration/numga_backend_b200/` (what a numga checkout installs as `numga/backend/b200/`):
classes that SUBCLASS the reference's `AbstractContext` / `AbstractMultiVector` / `AbstractConcreteOperator`
(numga/context.py:55-69, numga/operator/abstract.py:12-60) and execute through the numga_b200 engine.

CPU tier (build container only: needs /root/reference): the REFERENCE'S OWN TEST FUNCTIONS
  numga/backend/test/test_numpy.py::test_basic (lines 11-31), ::test_geometry_pga3d (269-276),
  numga/multivector/extension/test/test_logexp.py::test_bisect (11-34)
are imported unmodified and run with `NumpyContext` swapped for the adapter's context; kernels are executed by the
numpy IR interpreter (tests/cpu_emulation.py).  GPU tier: the same scenarios on cuda:0 against golden outputs of the
reference (tests/golden/adapter.npz, made by tests/golden/make_golden_adapter.py); no reference needed there.
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HAVE_REFERENCE = os.path.isdir('/root/reference/numga')


def _load_adapter():
    """numga must be importable: the reference itself where it exists (build container), else the adapter cannot load."""
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    from ref_loader import load_reference
    load_reference()
    if os.path.join(ROOT, 'integration') not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, 'integration'))
    import numga_backend_b200
    return numga_backend_b200


needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason='the reference (numga) is only present in the build container')


def _emulated_factory(A, dtype=torch.float64):
    """`NumpyContext(algebra, dtype=...)` look-alike that builds adapter contexts on CPU tensors (kernels emulated)."""
    from cpu_emulation import emulated_context

    def make(algebra, dtype=None, **kw):
        from numga.algebra.algebra import Algebra
        alg = algebra if isinstance(algebra, Algebra) else Algebra(algebra)
        engine = emulated_context(alg.description.description_str, torch.float64, mode='faithful')
        engine.lazy = bool(kw.get('lazy', False))
        return A.B200Context(alg, dtype=torch.float64, device='cpu', mode='faithful', engine=engine)
    return make


@needs_reference
def test_adapter_subclasses_the_reference_interfaces():
    A = _load_adapter()
    from numga.context import AbstractContext
    from numga.multivector.multivector import AbstractMultiVector
    from numga.operator.abstract import AbstractConcreteOperator
    assert issubclass(A.B200Context, AbstractContext)
    assert issubclass(A.B200MultiVector, AbstractMultiVector)
    assert issubclass(A.B200Operator, AbstractConcreteOperator)
    from cpu_emulation import emulate_launches
    with emulate_launches():
        ga = _emulated_factory(A)('x+y+z+w0')
        Q, V = ga.subspace.even_grade(), ga.subspace.vector()
        op = ga.operator.sandwich(Q, V)
        assert isinstance(op, A.B200Operator) and op is ga.operator.sandwich(Q, V)          # numga's operator cache
        q = ga.multivector.motor(np.arange(1, 9.0))
        v = ga.multivector.vector([1, 2, 3, 4.0])
        assert isinstance(q >> v, A.B200MultiVector) and (q >> v).subspace is V
        assert (q >> v).values.tolist() == [30, -60, -90, -116]
        assert (q * ga.multivector.motor(np.arange(8, 0, -1.0))).values.tolist() == [-44, 32, 12, 46, -72, 102, 36, 114]
        assert q.symmetric_reverse_product().values.tolist() == [30, -16]
        # every computing method is ONE engine kernel; numga's dispatch tables are untouched (other backends unaffected)
        x = 0.02 * np.arange(1, 31.0).reshape(5, 6)
        b = ga.multivector.bivector(x)
        n0 = len(ga.engine.engine.cache)
        m = b.exp()
        assert len(ga.engine.engine.cache) == n0 + 1 and m.subspace is Q
        from numga.backend.numpy.context import NumpyContext
        nb = NumpyContext('x+y+z+w0').multivector.bivector(x)
        # faithful mode: bit-exact on batches (numpy's einsum takes a BLAS path for a single unbatched multivector)
        assert np.array_equal(nb.exp().values, m.values.numpy())


@needs_reference
def test_reference_test_numpy_basic_and_geometry_with_the_context_swapped():
    A = _load_adapter()
    import numga.backend.test.test_numpy as T
    from cpu_emulation import emulate_launches
    saved = T.NumpyContext
    T.NumpyContext = _emulated_factory(A)
    try:
        with emulate_launches():
            T.test_basic()                      # numga/backend/test/test_numpy.py:11-31 (partial / sandwich_map / sandwich)
            T.test_geometry_pga3d()             # :269-276 (join, normalized, division, motor_square_root, sandwich)
    finally:
        T.NumpyContext = saved


@needs_reference
@pytest.mark.parametrize('descr', [(3, 0, 1), (2, 0, 1), (3, 0, 0), (4, 1, 0)])
def test_reference_test_logexp_bisect_with_the_context_swapped(descr):
    A = _load_adapter()
    import numga.multivector.extension.test.test_logexp as T
    from cpu_emulation import emulate_launches
    saved = T.NumpyContext
    T.NumpyContext = _emulated_factory(A)
    try:
        with emulate_launches():
            T.test_bisect(descr)                # numga/multivector/extension/test/test_logexp.py:11-34
    finally:
        T.NumpyContext = saved


@needs_reference
def test_reference_physics_example_runs_unchanged_on_the_adapter():
    """numga/examples/physics/{setup_chain,core,base}.py, imported from the reference and untouched, driven with the
    adapter context: set-up (point clouds, inertia maps, `.at[].set`, operator slicing) and one full sub-step
    reproduce the states the reference's numpy backend produced (tests/golden/physics.npz)."""
    A = _load_adapter()
    from numga.examples.physics.setup_chain import setup_bodies
    from cpu_emulation import emulate_launches
    G = np.load(os.path.join(HERE, 'golden', 'physics.npz'))
    with emulate_launches():
        ga = _emulated_factory(A)('x+y+z+w0')
        bodies, sets = setup_bodies(ga, n_bodies=12)
        assert np.allclose(bodies.inertia_inv.kernel.numpy(), G['f64/init/inertia_inv'], rtol=0, atol=1e-9)
        assert np.allclose(sets[0].anchors_map.kernel.numpy(), G['f64/set0/anchors_map'], rtol=0, atol=1e-14)
        bodies = bodies.copy(rate=ga.multivector.bivector(G['f64/start/rate']))
        new = bodies.integrate(1 / 200, sets)
        assert np.abs(new.motor.values.numpy() - G['f64/step1/motor']).max() < 1e-12
        assert np.abs(new.rate.values.numpy() - G['f64/step1/rate']).max() < 1e-10 * np.abs(G['f64/step1/rate']).max()


@needs_reference
def test_reference_physics_example_fuses_automatically_on_the_lazy_adapter():
    """The SAME unmodified reference files on `B200Context(..., lazy=True)`: multivector methods and operator calls are
    only recorded, `bodies.motor[idx]` / `inertia_inv[idx]` gathers and `.at[idx].set` scatters join the recorded
    expressions, and one sub-step of `Body.integrate` (numga/examples/physics/core.py:80-88; 174 array passes on the numpy
    backend) runs as a dozen generated kernels with the same result."""
    A = _load_adapter()
    from numga.examples.physics.setup_chain import setup_bodies
    from cpu_emulation import emulate_launches
    from numga_b200 import backend as B
    G = np.load(os.path.join(HERE, 'golden', 'physics.npz'))
    counter, launches = {'n': 0}, {}
    with emulate_launches():
        orig = B.CompiledFunction.launch               # (the emulated launch)

        def counting(self, *a, **k):
            counter['n'] += 1
            return orig(self, *a, **k)
        B.CompiledFunction.launch = counting
        try:
            for lazy in (False, True):
                ga = _emulated_factory(A)('x+y+z+w0', lazy=lazy)
                bodies, sets = setup_bodies(ga, n_bodies=12)
                bodies = bodies.copy(rate=ga.multivector.bivector(G['f64/start/rate']))
                bodies.rate.values, bodies.motor.values
                counter['n'] = 0
                new = bodies.integrate(1 / 200, sets)
                ga.realize(new.motor, new.rate)
                m, r = new.motor.values.numpy(), new.rate.values.numpy()
                launches[lazy] = counter['n']
                assert np.abs(m - G['f64/step1/motor']).max() < 1e-12
                assert np.abs(r - G['f64/step1/rate']).max() < 1e-10 * np.abs(G['f64/step1/rate']).max()
        finally:
            B.CompiledFunction.launch = orig
    assert launches[True] <= 16 < 100 < launches[False], launches
