"""Automatic expression fusion (`B200Context(lazy=True)`, numga_b200/lazy.py): chains written against
the unchanged multivector API are recorded and run as ONE kernel when their values are needed.  Results must equal the
eager path exactly (same traced arithmetic), the number of kernels must drop, and nothing may be computed twice."""
import contextlib
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'physics.npz'))
WHERE = [pytest.param('gpu', marks=pytest.mark.gpu), 'cpu']


@contextlib.contextmanager
def _contexts(descr, where, dtype=torch.float64):
    """(eager context, lazy context, launch counter)"""
    from numga_b200 import backend as B
    counter = {'n': 0}
    if where == 'gpu':
        make = lambda lazy: B.B200Context(descr, dtype=dtype, device='cuda:0', lazy=lazy)
        stack = contextlib.ExitStack()
    else:
        from cpu_emulation import emulate_launches, emulated_context
        stack = contextlib.ExitStack()
        stack.enter_context(emulate_launches())

        def make(lazy):
            ctx = emulated_context(descr, dtype, mode='fast')
            ctx.lazy = lazy
            return ctx
    with stack:
        orig = B.CompiledFunction.launch

        def counting(self, *a, **k):
            counter['n'] += 1
            return orig(self, *a, **k)
        B.CompiledFunction.launch = counting
        try:
            yield make(False), make(True), counter
        finally:
            B.CompiledFunction.launch = orig


def _np(mv):
    return mv.values.detach().cpu().numpy()


@pytest.mark.parametrize('where', WHERE)
def test_chain_becomes_one_kernel_with_identical_results(where):
    rng = np.random.default_rng(0)
    bv, pv = 0.4 * rng.standard_normal((7, 6)), rng.standard_normal((7, 4))
    with _contexts((3, 0, 1), where) as (eager, lazy, counter):
        def chain(ctx):
            b, p = ctx.multivector.bivector(bv), ctx.multivector.antivector(pv)
            m = b.exp()
            m = (m * m).normalized()
            return (m >> p) + p * 0.5 - 2, m.motor_log()
        counter['n'] = 0
        q0, l0 = chain(eager)
        q0, l0 = _np(q0), _np(l0)
        n_eager, counter['n'] = counter['n'], 0
        q1, l1 = chain(lazy)
        assert type(q1).__name__ == 'LazyMultiVector' and q1.pending and counter['n'] == 0
        assert q1.shape == (7,) and len(q1.subspace) == q0.shape[-1]
        lazy.realize(q1, l1)
        assert counter['n'] == 1 < n_eager
        assert np.allclose(_np(q1), q0, atol=1e-14) and np.allclose(_np(l1), l0, atol=1e-14)
        # the same expression again: no new kernel is compiled
        n_compiled = len(lazy.engine.cache)
        q2, l2 = chain(lazy)
        lazy.realize(q2, l2)
        assert len(lazy.engine.cache) == n_compiled and np.array_equal(_np(q2), _np(q1))


@pytest.mark.parametrize('where', WHERE)
def test_shared_intermediates_are_kept_dead_ones_are_not(where):
    rng = np.random.default_rng(1)
    with _contexts((3, 0, 1), where) as (eager, lazy, counter):
        x = lazy.multivector.bivector(0.3 * rng.standard_normal((5, 6)))
        a = x.exp()
        b, c = a * a, a + 1
        counter['n'] = 0
        b.values
        assert counter['n'] == 1 and not a.pending and c.pending        # `a` is still referenced: written by b's kernel
        c.values
        assert counter['n'] == 2
        d = (x.exp() * x.exp()).normalized()                             # anonymous intermediates stay in registers
        d.values
        assert len(list(lazy.engine.cache.values())[-1].out_subspaces) == 1


@pytest.mark.parametrize('where', WHERE)
def test_wide_bound_operators_are_re_evaluated_not_tabulated(where):
    """`points.inertia_map()` (numga/multivector/multivector.py:385-395; tabulated once per constraint set by the
    reference, examples/physics/core.py:133) under lazy=True: binding launches nothing, applying the map is ONE kernel
    that reads the 4-word points and the rates but no 36-entry table, and the map stays symbolic for the next use; a
    narrow map (a motor's 4x4 sandwich matrix from 8 words) is tabulated by the first kernel that uses it."""
    rng = np.random.default_rng(5)
    with _contexts((3, 0, 1), where) as (eager, lazy, counter):
        pts, rates = rng.standard_normal((7, 4)), rng.standard_normal((7, 6))
        want = _np(eager.multivector.antivector(pts).inertia_map()(eager.multivector.bivector(rates)))
        counter['n'] = 0
        imap = lazy.multivector.antivector(pts).inertia_map()
        assert counter['n'] == 0 and imap.shape == (7,)
        for k in range(2):                       # (a repeated call takes the C++ fast path, which `counter` does not see)
            got = imap(lazy.multivector.bivector(rates))
            got.values
            assert counter['n'] in (1, k + 1) and imap._flat is None and imap._lazy.pending
        assert sorted(len(s) for s in list(lazy.engine.cache.values())[-1].in_subspaces) == [4, 6]
        assert np.abs(_np(got) - want).max() <= 1e-14 * np.abs(want).max()
        assert np.array_equal(imap.kernel.cpu().numpy(), eager.multivector.antivector(pts).inertia_map().kernel.cpu().numpy())
        assert not imap._lazy.pending                                    # asking for the table computes it, once
        motors = lazy.multivector.motor(rng.standard_normal((7, 8)))
        smap = motors.sandwich_map()
        out = smap(lazy.multivector.vector(rng.standard_normal((7, 4))))
        out.values
        assert not smap._lazy.pending                                    # narrow map: written by the kernel that used it


@pytest.mark.parametrize('where', WHERE)
def test_reference_style_physics_step_fuses_automatically(where):
    """Body.integrate written like numga/examples/physics/core.py: ~170 kernels eagerly, 11 when lazy, same
    states as the unmodified reference."""
    from numga_b200.examples.physics import setup_bodies
    with _contexts('x+y+z+w0', where) as (eager, lazy, counter):
        launches = {}
        for name, ctx in (('eager', eager), ('lazy', lazy)):
            bodies, sets = setup_bodies(ctx, n_bodies=12)
            bodies = bodies.copy(rate=ctx.multivector.bivector(G['f64/start/rate']))
            counter['n'] = 0
            new = bodies.integrate(1 / 200, sets)
            m, r = _np(new.motor), _np(new.rate)
            launches[name] = counter['n']
            assert np.abs(m - G['f64/step1/motor']).max() < 1e-12 and np.abs(r - G['f64/step1/rate']).max() < 1e-10
        assert launches["lazy"] <= 12 < 100 < launches["eager"]      # 7 fused kernels + the 4 row copies `.at[].set` implies


@pytest.mark.parametrize('where', WHERE)
def test_pending_dag_is_bounded(where):
    with _contexts((3, 0, 0), where) as (eager, lazy, counter):
        lazy.lazy_limit = 16
        r = lazy.multivector.bivector(np.array([[0.1, 0.2, 0.3]])).exp()
        r.values
        m = r
        counter['n'] = 0
        for _ in range(40):
            m = m * r
        assert counter['n'] >= 2                     # flushed on the way: the pending DAG never exceeds the limit
        m.values
        rr = eager.multivector.bivector(np.array([[0.1, 0.2, 0.3]])).exp()
        ref = rr
        for _ in range(40):
            ref = ref * rr
        assert np.allclose(_np(m), _np(ref), atol=1e-12)


@pytest.mark.parametrize('where', WHERE)
def test_lazy_gathers_of_different_batch_shapes(where):
    """Replicated chains: per-constraint data gathered lazily over [c] items meets kernels over [2, c] items; a gather
    is done inside a kernel only when its index covers that kernel's whole batch, otherwise torch resolves it."""
    from numga_b200.examples.physics import replicate_chains, setup_bodies
    rng = np.random.default_rng(0)
    jitter = 0.3 * rng.standard_normal((600, 6))
    with _contexts('x+y+z+w0', where) as (eager, lazy, counter):
        states = {}
        for name, ctx in (('eager', eager), ('lazy', lazy)):
            bodies, sets = setup_bodies(ctx, n_bodies=12)
            big, bsets = replicate_chains(bodies, sets, 50)
            b = big.copy(rate=big.rate + ctx.multivector.bivector(jitter))
            for _ in range(2):
                b = b.integrate(1 / 200, bsets)
            states[name] = (_np(b.motor), _np(b.rate))
        assert np.abs(states['lazy'][0] - states['eager'][0]).max() < 1e-13
        assert np.abs(states['lazy'][1] - states['eager'][1]).max() < 1e-10


@pytest.mark.parametrize('where', WHERE)
def test_autograd_through_a_lazily_fused_chain(where):
    """A recorded chain with a `requires_grad` leaf: one fused forward kernel, one generated backward kernel, same
    gradient as the eager path."""
    rng = np.random.default_rng(0)
    b0, pv, qv = 0.3 * rng.standard_normal(6), rng.standard_normal((9, 4)), rng.standard_normal((9, 4))
    with _contexts((3, 0, 1), where) as (eager, lazy, counter):
        grads = {}
        for name, ctx in (('eager', eager), ('lazy', lazy)):
            b = torch.tensor(b0, dtype=torch.float64, device=ctx.device, requires_grad=True)
            p = ctx.multivector.antivector(pv)
            q = torch.tensor(qv, dtype=torch.float64, device=ctx.device)
            out = (ctx.multivector(values=b, subspace=ctx.subspace.bivector()).exp() >> p) + p * 0.5
            counter['n'] = 0
            loss = ((out.values - q) ** 2).sum()
            launches_fwd = counter['n']
            loss.backward()
            grads[name] = (b.grad.detach().cpu().numpy(), launches_fwd, counter['n'] - launches_fwd)
        assert np.allclose(grads['lazy'][0], grads['eager'][0], atol=1e-12)
        assert grads['lazy'][1] == 1 and grads['lazy'][2] == 1           # one forward kernel, one backward kernel


@pytest.mark.parametrize('which', ['test_bisect', 'test_motor_split', 'test_inverse_compare', 'test_study'])
def test_reference_identities_also_hold_under_lazy_fusion(which, monkeypatch):
    """The ported reference tests are written against the plain API; with NGB200_TEST_LAZY=1 the emulated contexts
    record and fuse everything, and the identities must still hold (every CPU-tier file can be rerun this way)."""
    import test_reference_identities as T
    monkeypatch.setenv('NGB200_TEST_LAZY', '1')
    descr = {'test_bisect': (3, 0, 1), 'test_motor_split': (3, 0, 1), 'test_inverse_compare': (2, 1, 0), 'test_study': (4, 0, 1)}[which]
    getattr(T, which)(descr, 'cpu')


@pytest.mark.gpu
def test_lazy_physics_step_captured_into_a_cuda_graph():
    """`ctx.capture` around the reference-style, lazily fused `Body.integrate`: the python recording runs once, every
    replay is one driver call; results equal the un-captured lazy step, and follow in-place updates of the state."""
    from numga_b200 import B200Context
    from numga_b200.examples.physics import replicate_chains, setup_bodies
    ctx = B200Context('x+y+z+w0', dtype=torch.float64, device='cuda:0', lazy=True)
    bodies, sets = setup_bodies(ctx, n_bodies=12)
    big, bsets = replicate_chains(bodies, sets, 50)
    big = big.copy(rate=big.rate + ctx.multivector.bivector(0.3 * np.random.default_rng(4).standard_normal((600, 6))))
    big.rate.values

    def advance(b):
        new = b.integrate(1 / 200, bsets)
        new.motor.values, new.rate.values
        return new
    want1 = advance(big)
    want2 = advance(want1)
    static = big.copy(motor=big.motor.copy(), rate=big.rate.copy())
    graph = ctx.capture(advance, static)
    assert graph.n_kernels == 11                 # 7 fused kernels + the 4 row copies `.at[].set` implies
    got1 = graph.replay()
    assert torch.equal(got1.motor.values, want1.motor.values) and torch.equal(got1.rate.values, want1.rate.values)
    static.motor.values.copy_(got1.motor.values)
    static.rate.values.copy_(got1.rate.values)
    got2 = graph.replay()
    assert torch.equal(got2.motor.values, want2.motor.values) and torch.equal(got2.rate.values, want2.rate.values)
