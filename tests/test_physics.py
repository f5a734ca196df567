"""Rigid-body physics core step (BASELINE.json config 5) against golden states produced by the
UNMODIFIED reference (`tests/golden/make_golden_physics.py` -> `tests/golden/physics.npz`).

CPU tier: the host orchestration (setup, gather/scatter, colour sets, the six fused kernels' traces) is checked
numerically by interpreting the traced IR with numpy (tests/cpu_emulation.py; test infrastructure, the product has no
CPU path).  GPU tier: the same checks with the real kernels through the C ABI.
"""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'physics.npz'))
DT = 1 / 200
TORCH = {'f32': torch.float32, 'f64': torch.float64}
# position-like state (motors) / velocity-like state (rates: finite differences divided by dt amplify rounding by 1/dt).
# Gates after k sub-steps come from tolerances.physics_gate (derived in profiles/r02_tolerances.md: fp32 rates 1.2e-3 /
# 1.6e-3 / 3e-3 after 1 / 2 / 5 sub-steps = 2x the measured kernel-vs-reference difference; the reference's own fp32
# rates are 5.5e-4 .. 8.4e-4 away from its fp64 result)
from tolerances import physics_gate
TOL = {'f64': physics_gate('f64', 1), 'f32': physics_gate('f32', 1)}


def _err(got, want):
    return float(np.abs(np.asarray(got) - want).max() / max(np.abs(want).max(), 1e-30))


def _check_setup(bodies, sets, tag):
    for f in ('motor', 'rate', 'first_moment', 'damping', 'gravity'):
        assert _err(getattr(bodies, f).values.cpu().numpy(), G[f'{tag}/init/{f}']) < (1e-6 if tag == 'f32' else 1e-14), f
    assert bodies.gravity.subspace.blades.astype(np.int64).tolist() == G[f'{tag}/init/gravity_blades'].tolist()
    assert _err(bodies.inertia.kernel.cpu().numpy(), G[f'{tag}/init/inertia']) < (1e-5 if tag == 'f32' else 1e-13)
    assert _err(bodies.inertia_inv.kernel.cpu().numpy(), G[f'{tag}/init/inertia_inv']) < (1e-4 if tag == 'f32' else 1e-11)
    for k, c in enumerate(sets):
        assert np.array_equal(c.body_idx.cpu().numpy(), G[f'{tag}/set{k}/body_idx'])
        assert _err(c.anchors.values.cpu().numpy(), G[f'{tag}/set{k}/anchors']) < (1e-6 if tag == 'f32' else 1e-14)
        assert _err(c.anchors_map.kernel.cpu().numpy(), G[f'{tag}/set{k}/anchors_map']) < (1e-6 if tag == 'f32' else 1e-14)
        assert _err(c.compliance.values.cpu().numpy(), G[f'{tag}/set{k}/compliance']) < 1e-6


def _stages(bodies, sets, tag):
    """The reference's sub-step, stage by stage, through the generic (one kernel per multivector method) path."""
    mt, rt_ = TOL[tag]
    new = bodies.pre_integrate(DT)
    assert _err(new.motor.values.cpu().numpy(), G[f'{tag}/pre/motor']) < mt
    assert _err(new.rate.values.cpu().numpy(), G[f'{tag}/pre/rate']) < mt
    for k, c in enumerate(sets):
        new = c.apply(new, dt=DT)
        assert _err(new.motor.values.cpu().numpy(), G[f'{tag}/apply{k}/motor']) < mt
    new = new.post_integrate(bodies, dt=DT)
    assert _err(new.motor.values.cpu().numpy(), G[f'{tag}/post/motor']) < mt
    assert _err(new.rate.values.cpu().numpy(), G[f'{tag}/post/rate']) < rt_
    for k, c in enumerate(sets):
        new = c.v_apply(new, dt=DT)
        assert _err(new.rate.values.cpu().numpy(), G[f'{tag}/vapply{k}/rate']) < rt_


def _fused_steps(bodies, sets, tag, steps=(1, 2, 5)):
    from numga_b200.examples.physics import FusedStep
    mt, rt_ = TOL[tag]
    fs = FusedStep(bodies, sets)
    b = bodies
    for step in range(1, max(steps) + 1):
        b = fs.integrate(b, DT)
        if step in steps:
            mt, rt_ = physics_gate(tag, step)
            assert _err(b.motor.values.cpu().numpy(), G[f'{tag}/step{step}/motor']) < mt
            assert _err(b.rate.values.cpu().numpy(), G[f'{tag}/step{step}/rate']) < rt_
    return b


def _chain_steps(bodies, sets, tag, steps=(1, 2, 5)):
    """The chain-resident single-kernel step (`ChainStep`): one sub-step per launch, then several per launch."""
    from numga_b200.examples.physics import ChainStep
    mt, rt_ = TOL[tag]
    cs = ChainStep(bodies, sets, unroll=1)
    b = bodies
    for step in range(1, max(steps) + 1):
        b = cs.integrate(b, DT)
        if step in steps:
            mt, rt_ = physics_gate(tag, step)
            assert _err(b.motor.values.cpu().numpy(), G[f'{tag}/step{step}/motor']) < mt
            assert _err(b.rate.values.cpu().numpy(), G[f'{tag}/step{step}/rate']) < rt_
    for unroll in sorted(set(steps) - {1}):                      # `unroll` sub-steps inside ONE launch
        b = ChainStep(bodies, sets, unroll=unroll).integrate(bodies, DT)
        mt, rt_ = physics_gate(tag, unroll)
        assert _err(b.motor.values.cpu().numpy(), G[f'{tag}/step{unroll}/motor']) < mt
        assert _err(b.rate.values.cpu().numpy(), G[f'{tag}/step{unroll}/rate']) < rt_
    return b


def _start(ctx, tag):
    from numga_b200.examples.physics import setup_bodies
    bodies, sets = setup_bodies(ctx, n_bodies=12)
    _check_setup(bodies, sets, tag)
    return bodies.copy(rate=ctx.multivector.bivector(G[f'{tag}/start/rate'])), sets


# ---- CPU tier (numpy interpretation of the traced IR) -------------------------------------------------------------
def test_physics_generic_and_fused_step_cpu_emulated():
    from cpu_emulation import emulate_launches, emulated_context
    from numga_b200.examples.physics import FusedStep, replicate_chains
    with emulate_launches():
        ctx = emulated_context('x+y+z+w0', torch.float64)
        bodies, sets = _start(ctx, 'f64')
        _stages(bodies, sets, 'f64')
        _fused_steps(bodies, sets, 'f64', steps=(1, 2))
        # the fused sub-step is six kernels for two colour sets; constraint kernels gather and scatter in place
        fs = FusedStep(bodies, sets)
        fs.integrate(bodies, DT)
        kinds = sorted({(k[0][1].__name__, v.program.in_mode.count(2), v.program.out_mode.count(2))
                       for k, v in ctx.engine.cache.items() if k[0][0] == "jit"})
        assert kinds == [('apply', 4, 2), ('post', 0, 0), ('pre', 0, 0), ('v_apply', 6, 2)]
        # replicated chains (config 5 workload): block-diagonal constraints, every chain evolves identically
        big, bsets = replicate_chains(bodies, sets, 3)
        one = fs.integrate(bodies, DT)
        many = FusedStep(big, bsets).integrate(big, DT)
        assert np.array_equal(many.motor.values.numpy().reshape(3, 12, 8), np.broadcast_to(one.motor.values.numpy(), (3, 12, 8)))
        assert np.array_equal(many.rate.values.numpy().reshape(3, 12, 6), np.broadcast_to(one.rate.values.numpy(), (3, 12, 6)))
        # the caller's state is never modified (immutable value semantics)
        assert _err(bodies.rate.values.numpy(), G['f64/start/rate']) == 0
        # the chain-resident pipeline kernel: same traced bodies, one launch per step / per several steps
        _chain_steps(bodies, sets, 'f64', steps=(1, 2))
        from numga_b200.examples.physics import ChainStep
        cs = ChainStep(big, bsets, unroll=2, bodies_per_tile=24)
        assert (cs.block, cs.blocks_per_tile) == (12, 2)          # 3 chains, 2 per tile: the last tile is partial
        many2 = cs.integrate(big, DT)
        two = fs.integrate(one, DT)
        assert np.array_equal(many2.motor.values.numpy().reshape(3, 12, 8), np.broadcast_to(two.motor.values.numpy(), (3, 12, 8)))
        assert np.array_equal(many2.rate.values.numpy().reshape(3, 12, 6), np.broadcast_to(two.rate.values.numpy(), (3, 12, 6)))
        assert _err(big.rate.values.numpy().reshape(3, 12, 6)[0], G['f64/start/rate']) == 0
        # exploit_zeros: the tables' exact zeros (inverse inertia of these bodies: 6 of 36 entries) are compiled out; same
        # states to rounding (sums lose their exact-zero terms only), far fewer instructions, and a guard against being
        # handed other tables than the ones inspected
        cz = ChainStep(big, bsets, unroll=2, bodies_per_tile=24, exploit_zeros=True)
        assert cz.compiled.program.info['n_instr'] < 0.9 * cs.compiled.program.info['n_instr']
        sparse2 = cz.integrate(big, DT)
        assert _err(sparse2.motor.values.numpy(), many2.motor.values.numpy()) < 1e-14
        assert _err(sparse2.rate.values.numpy(), many2.rate.values.numpy()) < 1e-12
        other = big.copy()
        other.inertia_inv = big.inertia_inv.at[0, 0, 0].set(1.0)
        with pytest.raises(ValueError):
            cz.integrate(other, DT)
        # `step` = the reference demo's main-loop step (run_chain_numpy.py:26-30): unroll sub-steps of dt / unroll
        from numga_b200.examples.physics import step
        stepped = step(bodies, sets, 2 * DT, unroll=2)
        assert np.array_equal(stepped.motor.values.numpy(), two.motor.values.numpy())
        assert np.array_equal(stepped.rate.values.numpy(), two.rate.values.numpy())
        with pytest.raises(ValueError):                           # a system without block structure is refused
            from numga_b200.examples.physics import _block_structure
            _block_structure(12, [torch.tensor([[0, 5], [11, 6]])], max_block=6)


def _pairs_against_default(ctx, tag, n_chains=5, bodies_per_tile=24):
    """ChainStep(pairs=True) -- two threads per constraint, one per side, coupling values exchanged by warp shuffle --
    against the one-thread-per-constraint ChainStep on perturbed replicated chains (a partial last tile included)."""
    from numga_b200.examples.physics import ChainStep, replicate_chains, setup_bodies
    bodies, sets = setup_bodies(ctx, n_bodies=12)
    big, bsets = replicate_chains(bodies, sets, n_chains)
    rng = np.random.default_rng(7)
    big = big.copy(rate=big.rate + ctx.multivector.bivector(0.3 * rng.standard_normal((12 * n_chains, 6))))
    ref = ChainStep(big, bsets, unroll=3, bodies_per_tile=bodies_per_tile).integrate(big, DT)
    cp = ChainStep(big, bsets, unroll=3, bodies_per_tile=bodies_per_tile, pairs=True)
    got = cp.integrate(big, DT)
    return cp, _np(ref.motor), _np(ref.rate), _np(got.motor), _np(got.rate)


def _np(mv):
    return mv.values.detach().cpu().numpy()


@pytest.mark.parametrize('mode', ['fast', 'faithful'])
def test_chain_step_with_two_threads_per_constraint_cpu_emulated(mode):
    from cpu_emulation import emulate_launches, emulated_context
    with emulate_launches():
        ctx = emulated_context('x+y+z+w0', torch.float64, mode=mode)
        cp, m0, r0, m1, r1 = _pairs_against_default(ctx, 'f64')
        assert cp.compiled.pipe.items_per_tile[1:] == [2 * 6 * 2, 2 * 5 * 2] and cp.compiled.pipe.threads_per_tile == 32
        assert np.array_equal(m0, m1) and np.array_equal(r0, r1)          # same operations in the same order, per side


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['f32', 'f64'])
@pytest.mark.parametrize('mode', ['fast', 'faithful'])
def test_chain_step_with_two_threads_per_constraint_gpu(tag, mode):
    from numga_b200 import B200Context
    ctx = B200Context('x+y+z+w0', dtype=torch.float32 if tag == 'f32' else torch.float64, mode=mode)
    for n_chains, bpt in ((5, 24), (70, None)):
        cp, m0, r0, m1, r1 = _pairs_against_default(ctx, tag, n_chains, bpt)
        if mode == 'faithful':
            assert np.array_equal(m0, m1) and np.array_equal(r0, r1)
        else:                                    # FMA contraction is chosen per program: equal to rounding
            mt, rt_ = physics_gate(tag, 3)
            assert _err(m1, m0) < mt and _err(r1, r0) < rt_


# ---- GPU tier ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('mode', ['fast', 'faithful'])
def test_physics_step_gpu_vs_reference(tag, mode):
    from numga_b200 import B200Context
    ctx = B200Context('x+y+z+w0', dtype=TORCH[tag], device='cuda:0', mode=mode)
    bodies, sets = _start(ctx, tag)
    _stages(bodies, sets, tag)
    _fused_steps(bodies, sets, tag)
    _chain_steps(bodies, sets, tag)


@pytest.mark.gpu
def test_physics_replicated_chains_gpu():
    """Config-5 workload shape: many independent chains in one batch; each chain must follow the single chain
    exactly (same kernels, same arithmetic), and the energy stays finite over a few sub-steps."""
    from numga_b200 import B200Context
    from numga_b200.examples.physics import FusedStep, replicate_chains
    ctx = B200Context('x+y+z+w0', dtype=torch.float32, device='cuda:0')
    bodies, sets = _start(ctx, 'f32')
    n_chains = 1000
    big, bsets = replicate_chains(bodies, sets, n_chains)
    fs1, fsn = FusedStep(bodies, sets), FusedStep(big, bsets)
    one, many = bodies, big
    for _ in range(3):
        one, many = fs1.integrate(one, DT), fsn.integrate(many, DT)
    m = many.motor.values.cpu().numpy().reshape(n_chains, 12, 8)
    r = many.rate.values.cpu().numpy().reshape(n_chains, 12, 6)
    assert np.array_equal(m, np.broadcast_to(one.motor.values.cpu().numpy(), m.shape))
    assert np.array_equal(r, np.broadcast_to(one.rate.values.cpu().numpy(), r.shape))
    assert np.isfinite(many.kinetic_energy().values.cpu().numpy()).all()
    # the chain-resident kernel on the same batch (1000 chains = 100 tiles of 10 chains; with 7 chains per tile the last
    # tile is partial): three sub-steps in one launch give the three-launch state of the six-kernel arrangement
    from numga_b200.examples.physics import ChainStep
    for bpt in (128, 84):
        got = ChainStep(big, bsets, unroll=3, bodies_per_tile=bpt).integrate(big, DT)
        assert np.array_equal(got.motor.values.cpu().numpy(), many.motor.values.cpu().numpy())     # same traced bodies,
        assert np.array_equal(got.rate.values.cpu().numpy(), many.rate.values.cpu().numpy())       # same arithmetic
        gm = got.motor.values.cpu().numpy().reshape(n_chains, 12, 8)
        assert np.array_equal(gm, np.broadcast_to(gm[0], gm.shape))          # every chain follows the same arithmetic


@pytest.mark.gpu
def test_chain_step_at_benchmark_scale_every_chain_follows_the_single_chain():
    """Size-independent property at (a quarter of) the BASELINE size: 349525 identical chains = 4.2 M bodies stepped by
    the chain-resident kernel (ten sub-steps in one launch) all equal the single chain stepped by the six-kernel
    arrangement, bit for bit, and stay finite."""
    from numga_b200 import B200Context
    from numga_b200.examples.physics import ChainStep, FusedStep, replicate_chains
    ctx = B200Context('x+y+z+w0', dtype=torch.float32, device='cuda:0')
    bodies, sets = _start(ctx, 'f32')
    n_chains = (1 << 22) // 12
    big, bsets = replicate_chains(bodies, sets, n_chains)
    got = ChainStep(big, bsets, unroll=10).integrate(big, DT)
    fs, one = FusedStep(bodies, sets), bodies
    for _ in range(10):
        one = fs.integrate(one, DT)
    m = got.motor.values.reshape(n_chains, 12, 8)
    r = got.rate.values.reshape(n_chains, 12, 6)
    assert bool(torch.isfinite(m).all()) and bool(torch.isfinite(r).all())
    assert bool((m == one.motor.values[None]).all()) and bool((r == one.rate.values[None]).all())


# ---- other signatures: the reference's chain demo runs in 2D elliptic space by default (run_chain_numpy.py:9-16) ------------
def _chain_vs_generic(ctx, descr, exact):
    from numga_b200.examples.physics import ChainStep, replicate_chains, setup_bodies
    bodies, sets = setup_bodies(ctx, n_bodies=6, compliance=1e-2 if descr == 'x+y+' else 1e-9)
    nb = bodies.rate.values.shape[-1]
    bodies = bodies.copy(rate=bodies.rate + ctx.multivector.bivector(0.3 * np.random.default_rng(1).standard_normal((6, nb))))
    ref = bodies.integrate(DT, sets).integrate(DT, sets)                 # written like the reference, one kernel per method
    big, bsets = replicate_chains(bodies, sets, 3)
    got = ChainStep(big, bsets, unroll=2, bodies_per_tile=12).integrate(big, DT)
    for f in ('motor', 'rate'):
        g = getattr(got, f).values.cpu().numpy().reshape(3, 6, -1)
        r = getattr(ref, f).values.cpu().numpy()
        assert np.isfinite(g).all()
        if exact:
            assert np.array_equal(g, np.broadcast_to(r, g.shape)), (descr, f)
        else:
            assert np.abs(g - r[None]).max() <= 1e-9 * max(1.0, np.abs(r).max()), (descr, f)


@pytest.mark.parametrize('descr', ['x+y+w0', 'x+y+z+', 'x+y+'])
def test_chain_step_in_other_signatures_cpu_emulated(descr):
    """2D PGA, 3D elliptic and the reference demo's default 2D elliptic space: the chain-resident kernel equals the
    reference-style `Body.integrate` (dimension- and signature-independent code, core.py docstring)."""
    from cpu_emulation import emulate_launches, emulated_context
    with emulate_launches():
        _chain_vs_generic(emulated_context(descr, torch.float64), descr, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize('descr', ['x+y+w0', 'x+y+z+', 'x+y+'])
def test_chain_step_in_other_signatures_gpu(descr):
    from numga_b200 import B200Context
    _chain_vs_generic(B200Context(descr, dtype=torch.float64, device='cuda:0'), descr, exact=False)


# ---- free tumbling (numga/examples/tennis_racket_theorem.py) ------------------------------------------------------------
def _tennis(ctx, n):
    from numga_b200.examples.physics import Body
    cube = 2 * ((np.arange(2 ** n)[:, None] & (1 << np.arange(n))) > 0) - 1
    points = ctx.multivector.vector(cube * (np.arange(n) + 1)).dual()
    nb = len(ctx.subspace.bivector())
    body = Body.from_point_cloud(points=points.repeat('... -> b ...', b=nb))
    body.rate = body.rate + ctx.multivector.bivector(np.eye(nb) + G[f'tennis{n}/jitter'])
    assert _err(body.kinetic_energy().values.cpu().numpy(), G[f'tennis{n}/energy0']) < 1e-12
    for step in range(1, 9):
        body = body.integrate(1 / 4)
        if step in (1, 8):
            assert _err(body.motor.values.cpu().numpy(), G[f'tennis{n}/step{step}/motor']) < 1e-9
            assert _err(body.rate.values.cpu().numpy(), G[f'tennis{n}/step{step}/rate']) < 1e-9
    assert _err(body.kinetic_energy().values.cpu().numpy(), G[f'tennis{n}/energy8']) < 1e-9


def test_tennis_racket_cpu_emulated():
    from cpu_emulation import emulate_launches, emulated_context
    with emulate_launches():
        _tennis(emulated_context((3, 0, 0), torch.float64), 3)


@pytest.mark.gpu
@pytest.mark.parametrize('n', [3, 4])
def test_tennis_racket_gpu(n):
    from numga_b200 import B200Context
    _tennis(B200Context((n, 0, 0), dtype=torch.float64, device='cuda:0'), n)
