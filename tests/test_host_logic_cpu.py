"""Host-side logic that needs no GPU: subspace predicates and dispatch, storage layout helpers, the
compile-only execution of the whole case catalogue (eager + fused), API surface."""
import numpy as np
import pytest
import torch

from cases import CASES
from numga_b200 import B200Context
from numga_b200.backend import flat_operand, soa_empty, to_soa


def test_dispatch_predicates_3dpga():
    ctx = B200Context((3, 0, 1), device='cpu')
    S = ctx.subspace
    M = S.even_grade()
    assert M.symmetric_reverse().named_str == '1,abcd' and M.symmetric_reverse().is_study
    assert S.vector().symmetric_reverse().equals.scalar()
    assert S.bivector().degenerate().is_degenerate and not S.bivector().is_degenerate
    assert S.translator().named_str == '1,ad,bd,cd' and S.translator().is_degenerate_scalar
    assert S.bireflection().named_str == '1,ab,ac,bc,ad,bd,cd'
    assert M.is_subalgebra and not S.vector().is_subalgebra
    assert S.vector().simplicity == 1 and M.simplicity == 2 and S.scalar().simplicity == 0
    assert (S.scalar().union(S.bivector())) is S.bireflection()       # flyweight identity
    assert S.xy_zw if False else S.ab_cd.named_str == 'ab,cd'


def test_soa_layout_and_flattening():
    v = soa_empty(8, (1000,), torch.float32, 'cpu')
    assert v.shape == (1000, 8) and v.stride() == (1, 1024)           # row pitch padded to 128 bytes
    t, ptr, cs, bs = flat_operand(v, 1000)
    assert (cs, bs) == (1024, 1) and ptr % 16 == 0
    w = soa_empty(4, (3, 5), torch.float32, 'cpu')
    assert w.shape == (3, 5, 4) and flat_operand(w, 15)[2:] == (15, 1)
    aos = torch.arange(40.).reshape(10, 4)
    s = to_soa(aos)
    assert torch.equal(s, aos) and s.stride() == (1, 32)
    sliced = s[::2]                                                     # strided batch still flattens
    assert flat_operand(sliced, 5)[2:] == (32, 2)
    single = torch.zeros(8)
    assert flat_operand(single, 1)[3] == 0


def test_coerce_array_converts_aos_to_soa():
    ctx = B200Context((3, 0, 1), device='cpu')
    m = ctx.multivector.motor(np.arange(80, dtype=np.float32).reshape(10, 8))
    assert m.values.shape == (10, 8) and m.values.stride(0) == 1 and m.shape == (10,)
    assert np.array_equal(m.numpy(), np.arange(80, dtype=np.float32).reshape(10, 8))
    unit = ctx.multivector.motor()
    assert unit.numpy().tolist() == [1, 0, 0, 0, 0, 0, 0, 0]
    with pytest.raises(AssertionError):
        ctx.multivector.motor(np.zeros((3, 7), np.float32))


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_whole_catalogue_runs_in_compile_only_mode(case):
    """Every case executes eagerly (one kernel per method) and fused (one kernel) with the right output
    subspace and batch shape; kernels are compiled for sm_100a, nothing is launched."""
    name, pqr, specs, scale, fn = case
    ctx = B200Context(pqr, device='cpu', compile_only=True)
    mvs = []
    for sub, kind in specs:
        s = getattr(ctx.subspace, sub)()
        mvs.append(ctx.multivector(values=torch.zeros((5, len(s)) if kind == 'batch' else (len(s),)), subspace=s))
    eager = fn(*mvs)
    fused = ctx.jit(fn)(*mvs)
    assert eager.subspace is fused.subspace
    assert eager.shape == fused.shape == (5,)
    assert eager.values.shape[-1] == len(eager.subspace)


def test_jit_cache_keys_on_subspace_and_broadcast_pattern():
    ctx = B200Context((3, 0, 1), device='cpu', compile_only=True)
    f = ctx.jit(lambda r, p: r >> p)
    R1, Rn = ctx.multivector.motor(torch.zeros(8)), ctx.multivector.motor(torch.zeros(6, 8))
    p = ctx.multivector.antivector(torch.zeros(6, 4))
    v = ctx.multivector.vector(torch.zeros(6, 4))
    f(R1, p); n1 = len(ctx.engine.cache)
    f(R1, p); assert len(ctx.engine.cache) == n1
    f(Rn, p); assert len(ctx.engine.cache) == n1 + 1
    f(R1, v); assert len(ctx.engine.cache) == n1 + 2
    scaled = ctx.jit(lambda m, s: m * s)
    scaled(Rn, 2.0); n2 = len(ctx.engine.cache)
    scaled(Rn, 3.5); assert len(ctx.engine.cache) == n2          # python scalars are runtime parameters


def test_multivector_data_movement_api():
    ctx = B200Context((3, 0, 1), device='cpu', compile_only=True)
    m = ctx.multivector.motor(torch.arange(48.).reshape(6, 8))
    assert m[2:4].shape == (2,) and len(m) == 6
    assert m.concatenate(m, axis=0).shape == (12,)
    assert m.reshape((2, 3)).shape == (2, 3) and m.reshape((2, 3)).flatten().shape == (6,)
    assert m.sum(axis=0).shape == ()
    upd = m.at[0].set(torch.ones(8))
    assert upd.values[0].tolist() == [1.0] * 8 and m.values[0, 1].item() == 1.0
    assert m.select.bivector().subspace is ctx.subspace.bivector()
    assert m.restrict[2].subspace is ctx.subspace.bivector()
    assert m.select[1].subspace is ctx.subspace.vector()
    assert (m + 1).subspace is m.subspace and (1 + m.select.bivector()).subspace is ctx.subspace.bireflection()
    assert ctx.multivector.ab.subspace.named_str == 'ab'


def test_jitted_functions_inline_into_each_other():
    """A jitted function called from inside another traced function becomes part of the caller's kernel."""
    ctx = B200Context((3, 0, 1), device='cpu', compile_only=True)
    S = ctx.subspace
    inner = ctx.jit(lambda m: m.normalized())
    outer = ctx.jit(lambda m, p: inner(m) >> p)
    ir_nested, _ = ctx.trace(lambda m, p: inner(m) >> p, S.even_grade(), S.antivector())
    ir_flat, _ = ctx.trace(lambda m, p: m.normalized() >> p, S.even_grade(), S.antivector())
    assert ir_nested['instr'] == ir_flat['instr']
    n = len(ctx.engine.cache)
    outer(ctx.multivector.motor(torch.zeros(4, 8)), ctx.multivector.antivector(torch.zeros(4, 4)))
    assert len(ctx.engine.cache) == n + 1            # one kernel for the whole thing, none for `inner` alone


def test_batch_axis_sums_and_row_gathers_are_generated_kernels():
    """`sum` / `mean` over a batch axis and `mv[index_array]` (numga/backend/numpy/multivector.py:30-56) run as generated
    kernels: slices added in registers, a batch-reduced output for the only batch axis, rows gathered straight into
    structure-of-arrays layout.  Semantics = numpy's on the values array (negative axes count the component axis)."""
    from cpu_emulation import emulate_launches, emulated_context
    rng = np.random.default_rng(3)
    with emulate_launches():
        ctx = emulated_context((3, 0, 1), torch.float64)
        x = rng.standard_normal((5, 7, 4))
        p = ctx.multivector.antivector(x)
        n0 = len(ctx.engine.cache)
        for axis in (0, 1, -2, -3):
            assert np.allclose(p.sum(axis).numpy(), x.sum(axis), rtol=0, atol=1e-14), axis
            assert np.allclose(p.mean(axis).numpy(), x.mean(axis), rtol=0, atol=1e-14), axis
            assert p.sum(axis).subspace is p.subspace and p.sum(axis).shape == x.sum(axis).shape[:-1]
        assert any(k[0][0] == 'sum_slices' for k in ctx.engine.cache) and len(ctx.engine.cache) > n0
        flat = ctx.multivector.antivector(x.reshape(35, 4))
        assert np.allclose(flat.sum(0).numpy(), x.reshape(35, 4).sum(0), rtol=0, atol=1e-14) and flat.sum(0).shape == ()
        assert np.allclose(flat.mean(-2).numpy(), x.reshape(35, 4).mean(0), rtol=0, atol=1e-14)
        assert any(k[0][0] == 'reduce' for k in ctx.engine.cache)
        with pytest.raises(AssertionError):
            p.sum(-1)
        big = ctx.multivector.antivector(rng.standard_normal((200, 3, 4)))         # more than 64 slices: folded
        assert np.allclose(big.sum(0).numpy(), big.numpy().sum(0), rtol=0, atol=1e-13)
        # row gathers: numpy index arrays, torch index tensors, negative and repeated indices, 2-D index arrays
        idx = np.array([3, 0, 34, -1, 3])
        got = flat[idx]
        assert np.array_equal(got.numpy(), x.reshape(35, 4)[idx]) and got.values.stride(0) == 1
        idx2 = torch.tensor([[0, 1, 2], [-3, 33, 5]])
        assert np.array_equal(flat[idx2].numpy(), x.reshape(35, 4)[idx2.numpy()]) and flat[idx2].shape == (2, 3)
        assert any(k[0] == ('gather',) for k in ctx.engine.cache)


def test_at_set_of_whole_rows_is_a_generated_copy_and_scatter():
    """`mv.at[index_array].set(rows)` returns a modified copy written by generated kernels (copy + INDEXED-output
    scatter); the original is untouched; other index forms keep torch semantics."""
    from cpu_emulation import emulate_launches, emulated_context
    rng = np.random.default_rng(4)
    with emulate_launches():
        ctx = emulated_context((3, 0, 1), torch.float64)
        x = rng.standard_normal((20, 8))
        m = ctx.multivector.motor(x)
        idx = np.array([[3, 7, 0], [19, -2, 5]])
        rows = ctx.multivector.motor(rng.standard_normal((2, 3, 8)))
        new = m.at[idx].set(rows)
        want = x.copy()
        want[idx] = rows.numpy()
        assert np.array_equal(new.numpy(), want) and np.array_equal(m.numpy(), x) and new.subspace is m.subspace
        assert any(k[0] == ('scatter_rows',) for k in ctx.engine.cache) and any(k[0] == ('copy_rows',) for k in ctx.engine.cache)
        one = m.at[4].set(ctx.multivector.motor(np.ones(8)))              # scalar index: torch path
        assert np.array_equal(one.numpy()[4], np.ones(8)) and np.array_equal(one.numpy()[5], x[5])


def test_pipeline_index_operands_are_bounds_checked_on_the_host():
    """The pipeline kernel trusts its index operands; `CompiledPipeline.validate` (run by `launch`) refuses rows that
    fall outside the addressing item's tile (shared buffers) or outside the array (global operands)."""
    from numga_b200 import B200Context
    from numga_b200.pipeline import Pipeline
    ctx = B200Context((3, 0, 1), device='cpu', compile_only=True)
    S = ctx.subspace
    pipe = Pipeline(ctx, {'node': 8, 'edge': 4})
    a, out = pipe.array('a', S.vector(), 'node'), pipe.array('out', S.vector(), 'edge')
    i0, i1 = pipe.index('i0', 'edge'), pipe.index('i1', 'edge')
    buf = pipe.buffer(S.vector(), 'node')
    pipe.phase('node', lambda x: x * 2, [a], [buf])
    pipe.phase('edge', lambda x, y: x - y, [buf[i0], buf[i1]], [out])
    cp = pipe.build()
    arrays = {'a': ctx.multivector.vector(torch.zeros(16, 4)), 'out': ctx.multivector.vector(torch.zeros(8, 4)),
              'i0': torch.tensor([0, 1, 2, 3, 8, 9, 10, 11]), 'i1': torch.tensor([7, 6, 5, 4, 15, 14, 13, 12])}
    cp.launch(arrays, {'node': 16, 'edge': 8})                               # compile-only: validated, not launched
    bad = dict(arrays, i1=torch.tensor([7, 6, 5, 4, 15, 14, 13, 3]))         # edge 7 (tile 1) reaches into tile 0
    with pytest.raises(ValueError, match="'i1'"):
        cp.launch(bad, {'node': 16, 'edge': 8})
    with pytest.raises(ValueError, match="'i0'"):
        cp.launch(dict(arrays, i0=torch.tensor([0, 1, 2, 3, 8, 9, 10, 16])), {'node': 16, 'edge': 8})


def test_pipeline_refuses_an_output_that_aliases_a_read_only_operand():
    from numga_b200 import B200Context
    from numga_b200.pipeline import Pipeline
    ctx = B200Context((3, 0, 1), device='cpu', compile_only=True)
    S = ctx.subspace
    pipe = Pipeline(ctx, {'node': 8})
    a, out = pipe.array('a', S.vector(), 'node'), pipe.array('out', S.vector(), 'node')
    pipe.phase('node', lambda x: x * 2, [a], [out])
    cp = pipe.build()
    x = ctx.multivector.vector(torch.zeros(16, 4))
    cp.launch({'a': x, 'out': ctx.multivector.vector(torch.zeros(16, 4))}, {'node': 16})
    with pytest.raises(ValueError, match='shares memory'):
        cp.launch({'a': x, 'out': x}, {'node': 16})
