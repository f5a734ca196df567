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
