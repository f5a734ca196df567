"""Edge cases of the non-linear chains (zero bivector, identity / null / tiny / huge motors, pure translators and
rotors, angles near pi) against outputs of the UNMODIFIED reference (`tests/golden/make_golden_edges.py` ->
`tests/golden/edges.npz`), including the rows where the reference itself returns nan.

  oracle                      bit-exact (pins the oracle on these inputs too)
  faithful programs           bit-exact on every row the reference keeps finite; rows where the reference overflows
                              to inf / nan must be non-finite here too.  (The reference's dense einsum multiplies
                              the ZERO kernel entries with inf intermediates, 0 * inf = nan, and so smears nan over
                              all components; a sparse kernel leaves them +-inf.  Same rows, different non-finite
                              pattern.)  CPU tier = numpy interpretation of the traced IR, GPU tier = real kernels
  fast programs (GPU)         within 1e-5 (fp32) / 1e-12 (fp64) wherever the reference is finite (fp32: except inputs whose
                              squared norm is subnormal, which the .ftz MUFU forms flush to zero)
"""
import os

import numpy as np
import pytest
import torch

from edge_cases import EDGE_CASES, PGA3


def _same_up_to_nonfinite_pattern(got, want):
    ok = np.isfinite(want).all(axis=-1)
    return np.array_equal(got[ok], want[ok]) and not np.isfinite(got[~ok]).all(axis=-1).any()

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'edges.npz'))
TORCH = {np.float32: torch.float32, np.float64: torch.float64}
IDS = [c[0] for c in EDGE_CASES]


def _golden(name, dtype):
    tag = f'{name}/{np.dtype(dtype).name}'
    return Z[f'{tag}/in'], Z[f'{tag}/out'], Z[f'{tag}/out_blades']


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('case', EDGE_CASES, ids=IDS)
def test_oracle_matches_reference_on_edge_cases(case, dtype):
    from oracle import numga_oracle as O
    name, sub, _, fn = case
    x, want, blades = _golden(name, dtype)
    with np.errstate(all='ignore'):
        got = fn(O.context(PGA3, dtype).multivector(sub, x))
    assert list(got.blades) == blades.tolist()
    assert np.array_equal(got.values, want, equal_nan=True)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('case', EDGE_CASES, ids=IDS)
def test_faithful_programs_on_edge_cases_cpu_emulated(case, dtype):
    from cpu_emulation import emulate_launches, emulated_context
    name, sub, _, fn = case
    x, want, blades = _golden(name, dtype)
    with emulate_launches():
        ctx = emulated_context(PGA3, TORCH[dtype], mode='faithful')
        got = ctx.jit(fn)(ctx.multivector(values=x, subspace=getattr(ctx.subspace, sub)()))
    assert got.subspace.blades.astype(np.int64).tolist() == blades.tolist()
    assert _same_up_to_nonfinite_pattern(got.values.numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize('path', ['eager', 'jit'])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('case', EDGE_CASES, ids=IDS)
def test_edge_cases_gpu(case, dtype, path):
    from numga_b200 import B200Context
    name, sub, _, fn = case
    x, want, blades = _golden(name, dtype)
    # faithful: bit-exact, nan where the reference has nan
    ctx = B200Context(PGA3, dtype=TORCH[dtype], device='cuda:0', mode='faithful')
    f = ctx.jit(fn) if path == 'jit' else fn
    got = f(ctx.multivector(values=x, subspace=getattr(ctx.subspace, sub)()))
    assert got.subspace.blades.astype(np.int64).tolist() == blades.tolist()
    assert _same_up_to_nonfinite_pattern(got.numpy(), want)
    # fast: within tolerance on every row where the reference is finite
    ctx = B200Context(PGA3, dtype=TORCH[dtype], device='cuda:0', mode='fast')
    f = ctx.jit(fn) if path == 'jit' else fn
    fast = f(ctx.multivector(values=x, subspace=getattr(ctx.subspace, sub)())).numpy()
    ok = np.isfinite(want).all(axis=-1)
    if dtype == np.float32:
        # FAST fp32 contract: rcp / sqrt / rsqrt flush subnormal arguments to zero (MUFU .ftz forms); a
        # motor of magnitude 1e-20 has a subnormal squared norm (1e-40) and is outside that contract
        ok &= np.abs(x).max(axis=-1) > 1e-18
    # measured on B200 (scripts/edge_tolerances.py): worst case 1.0e-6 (fp32, roundtrip) / 7.5e-16 (fp64) on these
    # fixtures, so the north-star gates apply to the edge cases too (round 1 used 5e-3 / 1e-8 for the chains)
    tol = {np.float32: 1e-5, np.float64: 1e-12}[dtype]
    scale = np.maximum(np.abs(want[ok]).max(axis=-1, keepdims=True), 1e-30)
    assert np.all(np.abs(fast[ok] - want[ok]) <= tol * scale), float((np.abs(fast[ok] - want[ok]) / scale).max())
