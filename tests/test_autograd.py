"""torch.autograd through the generated kernels: the backward of every compiled function is
ANOTHER generated kernel (trace.build_vjp), checked here against central finite differences of the forward pass.

CPU tier: forward and backward IR interpreted with numpy (tests/cpu_emulation.py).  GPU tier: the real kernels."""
import contextlib

import numpy as np
import pytest
import torch


@contextlib.contextmanager
def _context(descr, where, dtype=torch.float64):
    if where == 'gpu':
        from numga_b200 import B200Context
        yield B200Context(descr, dtype=dtype, device='cuda:0')
    else:
        from cpu_emulation import emulate_launches, emulated_context
        with emulate_launches():
            yield emulated_context(descr, dtype, mode='fast')


WHERE = [pytest.param('gpu', marks=pytest.mark.gpu), 'cpu']

CASES = {
    'motor_product': ((3, 0, 1), ['even_grade', 'even_grade'], [(7,), (7,)], lambda a, b: a * b),
    'sandwich_broadcast_motor': ((3, 0, 1), ['even_grade', 'antivector'], [(), (9,)], lambda r, p: r >> p),
    'normalize_then_apply': ((3, 0, 1), ['even_grade', 'antivector'], [(), (6,)], lambda r, p: r.normalized() >> p),
    'exp_log_chain': ((3, 0, 1), ['bivector'], [(5,)], lambda b: (b.exp() * b.exp()).normalized().motor_log()),
    'partial_broadcast': ((3, 0, 1), ['even_grade', 'even_grade'], [(2, 3), (3,)], lambda a, b: a.reverse_product(b)),
    'cga_wedge_inner': ((4, 1, 0), ['vector', 'bivector'], [(4,), (4,)], lambda v, b: (v ^ b) + (v | b)),
    'norm_and_divide': ((3, 0, 0), ['even_grade', 'vector'], [(5,), (5,)], lambda r, v: (r >> v) / r.norm()),
}


def _loss(ctx, fn, subs, tensors, weights):
    out = fn(*[ctx.multivector(values=t, subspace=getattr(ctx.subspace, s)()) for s, t in zip(subs, tensors)])
    outs = out if isinstance(out, tuple) else (out,)
    return sum((o.values * w).sum() for o, w in zip(outs, weights)), outs


@pytest.mark.parametrize('where', WHERE)
@pytest.mark.parametrize('path', ['eager', 'jit'])
@pytest.mark.parametrize('name', list(CASES))
def test_gradients_match_finite_differences(name, path, where):
    descr, subs, shapes, fn = CASES[name]
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    with _context(descr, where) as ctx:
        f = ctx.jit(fn) if path == 'jit' else fn
        widths = [len(getattr(ctx.subspace, s)()) for s in subs]
        xs = [torch.tensor(0.5 * rng.standard_normal(sh + (w,)), dtype=torch.float64, device=ctx.device, requires_grad=True)
              for sh, w in zip(shapes, widths)]
        with torch.no_grad():
            _, outs = _loss(ctx, f, subs, xs, [1.0])
        weights = [torch.tensor(rng.standard_normal(tuple(o.values.shape)), dtype=torch.float64, device=ctx.device) for o in outs]
        loss, _ = _loss(ctx, f, subs, xs, weights)
        loss.backward()
        for k, x in enumerate(xs):
            assert x.grad is not None and tuple(x.grad.shape) == tuple(x.shape)
            num = torch.zeros_like(x)
            flat, nflat = x.detach().clone().reshape(-1), num.reshape(-1)
            for i in range(flat.numel()):
                vals = []
                for h in (+1e-6, -1e-6):
                    y = flat.clone(); y[i] += h
                    ys = [t.detach() for t in xs]; ys[k] = y.reshape(x.shape)
                    with torch.no_grad():
                        vals.append(float(_loss(ctx, f, subs, ys, weights)[0]))
                nflat[i] = (vals[0] - vals[1]) / 2e-6
            err = float((x.grad - num).abs().max() / max(1.0, float(num.abs().max())))
            assert err < 1e-6, (name, k, err)


@pytest.mark.parametrize('where', WHERE)
def test_gradient_descent_recovers_a_motor(where):
    """The use the reference's readme motivates (differentiable pipelines): fit a motor to point correspondences."""
    rng = np.random.default_rng(3)
    with _context((3, 0, 1), where) as ctx:
        S = ctx.subspace
        target = ctx.multivector.bivector(0.3 * rng.standard_normal(6)).exp()
        p = ctx.multivector.antivector(np.concatenate([rng.standard_normal((32, 3)), np.ones((32, 1))], axis=1))
        q = (target >> p).values.detach()
        b = torch.zeros(6, dtype=torch.float64, device=ctx.device, requires_grad=True)
        opt = torch.optim.Adam([b], lr=0.05)
        first = None
        for _ in range(60 if where == 'gpu' else 25):
            opt.zero_grad()
            m = ctx.multivector(values=b, subspace=S.bivector()).exp()
            loss = (((m >> p).values - q) ** 2).sum()
            loss.backward()
            opt.step()
            first = loss.item() if first is None else first
        assert loss.item() < 0.05 * first


def test_vjp_program_of_a_bilinear_operator_is_the_transposed_table():
    """Backward of out = K(a, b) w.r.t. a is K contracted with (cotangent, b): same nnz, no extra arithmetic."""
    from numga_b200 import B200Context, runtime as rt
    from numga_b200.trace import build_vjp
    ctx = B200Context((3, 0, 1), dtype=torch.float32, device='cpu', compile_only=True)
    S = ctx.subspace
    ir, _ = ctx.trace(lambda a, b: a * b, S.even_grade(), S.even_grade())
    vjp = build_vjp(ir, np.float32)
    assert vjp['in_width'] == [8, 8, 8] and vjp['out_width'] == [8, 8]
    muls = sum(1 for op, *_ in vjp['instr'] if op == rt.OP_MUL)
    assert muls <= 2 * 64 + 16        # two transposed 64-term tables (+ sharing slack), forward products are dead code


@pytest.mark.parametrize('where', WHERE)
def test_broadcast_operand_gradient_is_reduced_inside_the_kernel(where):
    """d loss / d R for ONE motor applied to a large batch: the backward kernel sums the per-item contributions itself
    (NGB200_REDUCED output: shuffle + shared + one atomicAdd per CTA) instead of writing batch x 8 values."""
    from numga_b200 import runtime as rt
    rng = np.random.default_rng(9)
    n = (1 << 18) + 77 if where == 'gpu' else 300
    with _context((3, 0, 1), where) as ctx:
        S = ctx.subspace
        R = torch.tensor(rng.standard_normal(8), dtype=torch.float64, device=ctx.device, requires_grad=True)
        p = ctx.multivector.antivector(rng.standard_normal((n, 4)))
        w = torch.tensor(rng.standard_normal((n, 4)), dtype=torch.float64, device=ctx.device)
        f = ctx.jit(lambda r, x: r >> x)
        loss = (f(ctx.multivector(values=R, subspace=S.even_grade()), p).values * w).sum()
        loss.backward()
        compiled = [v for k, v in ctx.engine.cache.items() if k[0][0] == 'jit'][0]
        assert compiled.vjp((0,)).program.out_mode == (rt.REDUCED,)      # only R needs a gradient: nothing is written for p
        # reference: the same gradient through per-item motors (no broadcasting, no reduction in the kernel)
        Rn = R.detach().expand(n, 8).clone().requires_grad_(True)
        loss2 = (f(ctx.multivector(values=Rn, subspace=S.even_grade()), p).values * w).sum()
        loss2.backward()
        want = Rn.grad.sum(dim=0)
        assert tuple(R.grad.shape) == (8,)
        assert float((R.grad - want).abs().max() / want.abs().max()) < 1e-10


def test_vjp_of_ir_with_duplicate_values_and_abs_at_zero():
    """Regressions (round-1 ADVICE): (1) forward IR that is NOT hash-consed -- two identical instructions that the
    replay merges into one value -- must keep every consumer's adjoint contribution (the reverse sweep runs over the
    new ids in descending order); (2) d|x|/dx at x == 0 is 0, as in torch, not -1."""
    from ir_interp import run_ir
    from numga_b200 import runtime as rt
    from numga_b200.trace import build_vjp
    ir = dict(in_width=[1, 1], in_mode=[rt.VARYING, rt.VARYING], out_width=[1], n_params=0, consts=[],
              instr=[(rt.OP_INPUT, 0, 0, 0), (rt.OP_INPUT, 1, 0, 0), (rt.OP_MUL, 0, 1, 0), (rt.OP_SQRT, 2, 0, 0),
                     (rt.OP_MUL, 0, 1, 0), (rt.OP_ADD, 3, 4, 0)],          # sqrt(a*b) + a*b, the product stated twice
              stores=[(0, 0, 5)])
    vjp = build_vjp(ir, np.float64)
    a, b, g = np.array([[1.5], [0.25]]), np.array([[2.0], [4.0]]), np.array([[1.0], [3.0]])
    ga, gb = run_ir(vjp, [a, b, g], np.float64)
    want = g * (0.5 / np.sqrt(a * b) + 1.0)
    assert np.allclose(ga, want * b, rtol=1e-14) and np.allclose(gb, want * a, rtol=1e-14)
    ir_abs = dict(in_width=[1], in_mode=[rt.VARYING], out_width=[1], n_params=0, consts=[],
                  instr=[(rt.OP_INPUT, 0, 0, 0), (rt.OP_ABS, 0, 0, 0)], stores=[(0, 0, 1)])
    x, g = np.array([[-2.0], [0.0], [3.0]]), np.array([[5.0], [5.0], [5.0]])
    (gx,) = run_ir(build_vjp(ir_abs, np.float64), [x, g], np.float64)
    assert gx.reshape(-1).tolist() == [-5.0, 0.0, 5.0]
    t = torch.tensor(x.reshape(-1), requires_grad=True)
    t.abs().backward(torch.tensor(g.reshape(-1)))
    assert t.grad.tolist() == gx.reshape(-1).tolist()


def test_vjp_of_atan2_matches_torch():
    """The ATAN2 opcode (closed-form motor logarithm, numga_b200/optimized.py) and its adjoint, all four quadrants."""
    from ir_interp import run_ir
    from numga_b200 import runtime as rt
    from numga_b200.trace import build_vjp
    ir = dict(in_width=[1, 1], in_mode=[rt.VARYING, rt.VARYING], out_width=[1], n_params=0, consts=[],
              instr=[(rt.OP_INPUT, 0, 0, 0), (rt.OP_INPUT, 1, 0, 0), (rt.OP_ATAN2, 0, 1, 0)], stores=[(0, 0, 2)])
    y, x = np.array([[0.5], [0.5], [-2.0], [-2.0]]), np.array([[1.5], [-1.5], [0.25], [-0.25]])
    g = np.array([[1.0], [2.0], [3.0], [4.0]])
    (out,) = run_ir(ir, [y, x], np.float64)
    gy, gx = run_ir(build_vjp(ir, np.float64), [y, x, g], np.float64)
    ty, tx = torch.tensor(y, requires_grad=True), torch.tensor(x, requires_grad=True)
    tout = torch.atan2(ty, tx)
    tout.backward(torch.tensor(g))
    assert np.allclose(out, tout.detach().numpy(), rtol=1e-15)
    assert np.allclose(gy, ty.grad.numpy(), rtol=1e-14) and np.allclose(gx, tx.grad.numpy(), rtol=1e-14)


@pytest.mark.parametrize('where', WHERE)
def test_torch_compile_fullgraph_motor_fitting(where):
    """`ctx.torch_op`: generated programs registered as `torch.library.custom_op`s (fake kernels from the IR's output
    layout, backward = the VJP program) survive `torch.compile(fullgraph=True)`: the motor-fitting loop of
    `test_gradient_descent_recovers_a_motor`, with loss AND gradient computed by ONE compiled graph."""
    rng = np.random.default_rng(3)
    with _context((3, 0, 1), where) as ctx:
        S = ctx.subspace
        target = ctx.multivector.bivector(0.3 * rng.standard_normal(6)).exp()
        p = ctx.multivector.antivector(np.concatenate([rng.standard_normal((32, 3)), np.ones((32, 1))], axis=1))
        q = (target >> p).values.detach()
        spin = ctx.torch_op(lambda b, x: b.exp() >> x, S.bivector(), S.antivector())

        def loss_fn(b, x, y):
            return ((spin(b, x) - y) ** 2).sum()
        # eager custom op: same value and gradient as the autograd.Function path
        b0 = torch.tensor(0.1 * rng.standard_normal(6), dtype=torch.float64, device=ctx.device, requires_grad=True)
        l_op = loss_fn(b0, p.values, q)
        (g_op,) = torch.autograd.grad(l_op, b0)
        m = ctx.multivector(values=b0, subspace=S.bivector()).exp()
        l_ref = (((m >> p).values - q) ** 2).sum()
        (g_ref,) = torch.autograd.grad(l_ref, b0)
        assert abs(l_op.item() - l_ref.item()) < 1e-12 * max(1, abs(l_ref.item()))
        assert float((g_op - g_ref).abs().max()) < 1e-10 * float(g_ref.abs().max())
        # compiled: fullgraph=True fails on any graph break (unsupported python, missing fake kernel ...)
        compiled = torch.compile(loss_fn, fullgraph=True, backend='inductor' if where == 'gpu' else 'aot_eager')
        b = torch.zeros(6, dtype=torch.float64, device=ctx.device, requires_grad=True)
        opt = torch.optim.Adam([b], lr=0.05)
        first = None
        for _ in range(60 if where == 'gpu' else 25):
            opt.zero_grad()
            loss = compiled(b, p.values, q)
            loss.backward()
            opt.step()
            first = loss.item() if first is None else first
        assert loss.item() < 0.05 * first
        torch.library.opcheck(spin.forward, ([b0.detach(), p.values],), test_utils=('test_schema', 'test_faketensor'))
