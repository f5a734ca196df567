"""Opt-in closed forms for 3D PGA (`numga_b200.optimized`), the counterpart of the reference's optional
`numga/multivector/extension/optimized.py`; tests ported from `extension/test/test_optimized.py:9-52`
(closed-form normalize / exp == the generic algorithms, atol 1e-8) plus the dispatch bookkeeping."""
import contextlib

import numpy as np
import pytest
import torch

from refutil import assert_close, random_non_motor, random_subspace, values_of


@contextlib.contextmanager
def _context(where):
    if where == 'gpu':
        from numga_b200 import B200Context
        yield B200Context('x+y+z+w0', dtype=torch.float64, device='cuda:0')
    else:
        from cpu_emulation import emulate_launches, emulated_context
        with emulate_launches():
            yield emulated_context('x+y+z+w0', torch.float64, mode='fast')


@pytest.fixture
def optimized():
    import numga_b200.optimized as opt
    yield opt
    opt.disable()


WHERE = [pytest.param('gpu', marks=pytest.mark.gpu), 'cpu']


@pytest.mark.parametrize('where', WHERE)
def test_normalize(where, optimized):
    rng = np.random.RandomState(0)
    with _context(where) as ga:
        m = random_non_motor(ga, (10,), rng=rng)
        n1 = m.normalized()
        optimized.enable()
        n2 = m.normalized()
        assert_close(n1, n2)
        assert_close(n2 * ~n2, 1)


@pytest.mark.parametrize('where', WHERE)
def test_exp(where, optimized):
    rng = np.random.RandomState(0)
    with _context(where) as ga:
        B = ga.subspace.bivector()
        for s in (B, B.degenerate(), B.nondegenerate()):
            for shape in ((10, 10), ()):
                b = random_subspace(ga, s, shape, rng)
                m1 = b.exp_bisect()
                optimized.enable()
                m2 = b.exp()
                optimized.disable()
                assert m1.subspace.blades.tolist() == m2.subspace.blades.tolist() or len(m2.subspace) <= len(m1.subspace)
                assert_close(m1, m2.select_subspace(m1.subspace), atol=1e-8)


@pytest.mark.parametrize('where', WHERE)
def test_closed_form_exp_is_one_small_kernel_and_handles_pure_translations(where, optimized):
    with _context(where) as ga:
        optimized.enable()
        B = ga.subspace.bivector()
        f = ga.jit(lambda b: b.exp())
        z = ga.multivector(B, np.array([[0, 0, 0, 1., 2., 3.], [0, 0, 0, 0, 0, 0], [1e-12, 0, 0, 0.5, 0, 0]]))
        got = values_of(f(z))
        assert np.allclose(got[0], [1, 0, 0, 0, 1, 2, 3, 0]) and np.allclose(got[1], [1, 0, 0, 0, 0, 0, 0, 0])
        assert np.isfinite(got).all()
        n_closed = len([c for k, c in ga.engine.cache.items() if k[0][0] == 'jit'][0].ir['instr'])
        optimized.disable()
        g = ga.jit(lambda b: b.exp())
        g(z)
        n_generic = len([c for k, c in ga.engine.cache.items() if k[0][0] == 'jit'][0].ir['instr'])
        assert n_closed < 60 < 400 < n_generic


@pytest.mark.parametrize('where', WHERE)
def test_closed_form_motor_log_equals_the_bisection_logarithm(where, optimized):
    """`log_motor` (ours; the reference's log is the bisection, logexp.py:101-117) against `motor_log_bisect` in fp64 and
    as the inverse of the closed-form exp: general motors, pure rotors, pure translators, the identity, a tiny angle."""
    rng = np.random.RandomState(3)
    with _context(where) as ga:
        B = ga.subspace.bivector()
        b = random_subspace(ga, B, (200,), rng) * 0.5
        special = np.zeros((5, 6))
        special[0, :3] = [0.3, -0.2, 0.9]              # pure rotation
        special[1, 3:] = [1.0, 2.0, -3.0]              # pure translation
        special[3] = [1e-9, 0, 0, 0.5, 0.25, -1.0]     # rotation angle below every threshold
        special[4] = [2e-2, 1e-2, -3e-2, 0.5, 0.25, -1.0]   # inside the series branch
        b = b.concatenate(ga.multivector(B, special), axis=0)
        m = b.exp_bisect().normalized()
        want = values_of(m.motor_log_bisect())
        optimized.enable()
        got = values_of(m.motor_log())
        f = ga.jit(lambda x: x.exp().normalized().motor_log())
        back = values_of(f(b))
        n_closed = len([c for k, c in ga.engine.cache.items() if k[0][0] == 'jit'][0].ir['instr'])
        optimized.disable()
        assert np.abs(got - want).max() < 1e-8 and np.abs(got - values_of(b)).max() < 1e-8
        assert np.abs(back - values_of(b)).max() < 1e-12          # closed-form round trip, one kernel
        assert n_closed < 250


@pytest.mark.parametrize('where', WHERE)
def test_gradient_of_the_closed_form_logarithm(where, optimized):
    """atan2 (the one opcode the logarithm adds) differentiates: d/db |log(exp(b))|^2 = 2 b."""
    rng = np.random.RandomState(4)
    with _context(where) as ga:
        if where != 'gpu':
            pytest.skip('autograd runs the VJP kernel on the device')
        optimized.enable()
        v = torch.tensor(0.4 * rng.randn(16, 6), device='cuda:0', dtype=torch.float64, requires_grad=True)
        b = ga.multivector(ga.subspace.bivector(), v)
        out = ga.jit(lambda x: x.exp().normalized().motor_log())(b)
        (out.values ** 2).sum().backward()
        assert torch.allclose(v.grad, 2 * v.detach(), atol=1e-9)


def test_enable_before_any_context_exists():
    """`opt.enable()` as the first statement of a program (no context yet: nothing to invalidate) must not raise."""
    import subprocess
    import sys
    code = ("import numga_b200.optimized as opt; opt.enable(); from numga_b200 import B200Context; "
            "c = B200Context((3, 0, 1), device='cpu', compile_only=True); opt.disable(); print('ok')")
    root = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=root)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stderr[-2000:]


def test_enable_disable_leave_other_algebras_alone(optimized):
    from numga_b200 import B200Context
    optimized.enable()
    vga = B200Context((3, 0, 0), dtype=torch.float64, device='cpu', compile_only=True)
    ir, _ = vga.trace(lambda b: b.exp(), vga.subspace.bivector())
    assert len(ir['instr']) > 100                 # (3,0,0) keeps the generic bisection
    pga = B200Context((3, 0, 1), dtype=torch.float64, device='cpu', compile_only=True)
    ir, _ = pga.trace(lambda b: b.exp(), pga.subspace.bivector())
    assert len(ir['instr']) < 60
    optimized.disable()
    ir, _ = pga.trace(lambda b: b.exp(), pga.subspace.bivector())
    assert len(ir['instr']) > 400
