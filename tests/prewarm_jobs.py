"""Build-time pre-compilation of the kernels the TEST-SUITE needs (case catalogue in both dtypes and both arithmetic
modes, the ported reference identity tests), so that `pytest -m gpu` on the GPU box finds every cubin in the cache.
A job provider for `numga_b200.prewarm.prewarm(extra=[(tests_dir, 'prewarm_jobs')])`, called by `__graft_entry__.build()`:
the dependency points from tests/ to the product, never the other way round."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def jobs():
    from cases import CASES
    import test_reference_identities as T
    out = [('case', c[0]) for c in CASES]
    for k, (fn, params) in enumerate(T.GPU_MATRIX):
        out += [('identity', (k, descr)) for descr in params]
    out += [('identity_single', k) for k in range(len(T.GPU_SINGLES))]
    out += [('physics_signature', d) for d in ('x+y+w0', 'x+y+z+', 'x+y+')]
    return out


def weight(job):
    """Biggest first: the 5-dimensional inverse sweeps dominate the wall clock."""
    return sum(job[1][1]) if job[0] == 'identity' else 0


def run_job(job):
    import torch
    from numga_b200.prewarm import _ctx_factory, _run
    kind, arg = job
    ctx_of = _ctx_factory()
    if kind == 'case':
        from cases import CASE_BY_NAME
        name, pqr, specs, scale, fn = CASE_BY_NAME[arg]
        for dtype in (torch.float32, torch.float64):
            for mode in ('fast', 'faithful'):
                ctx = ctx_of(pqr, dtype, mode)
                _run(ctx, fn, [(getattr(ctx.subspace, s)(), 'batch' if k == 'batch' else 'single') for s, k in specs])
    elif kind == 'physics_signature':
        import test_physics as TP
        try:
            TP._chain_vs_generic(ctx_of(arg, torch.float64, 'fast'), arg, exact=False)
        except AssertionError:
            pass
    else:
        import test_reference_identities as T
        try:
            if kind == 'identity':
                T.GPU_MATRIX[arg[0]][0](arg[1], 'compile')
            else:
                T.GPU_SINGLES[arg]('compile')
        except AssertionError:
            pass        # value assertions mean nothing on uninitialised memory; the kernels up to there are compiled
    return job
