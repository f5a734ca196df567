"""Compile (NVRTC, sm_100a) the generated kernels of the standard programs into the in-tree cubin cache.

Runs without a GPU: contexts are created in compile-only mode and the expressions of bench.py, of the physics example
and of the gather / scatter paths are executed on uninitialised host tensors, so that every kernel they need --
per-method eager kernels, fused `jit` kernels, operator-table kernels, pipeline kernels -- is generated, compiled and
stored under numga_b200/_kcache/.  On the GPU box these are cache hits (no JIT there).  The work is a list of
independent jobs spread over a process pool (the cache is written atomically, so workers never collide).

The test-suite's own programs (case catalogue, ported identity tests) are pre-compiled by `tests/prewarm_jobs.py`, which
plugs into the same pool through `prewarm(extra=...)`; this module does not import anything from tests/.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _paths():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)


def bench_programs(ctx_of):
    """The fused programs bench.py times, fast and faithful (kept in one place so build() and bench agree)."""
    import torch
    out = []
    for mode in ('fast', 'faithful'):
        pga32 = ctx_of((3, 0, 1), torch.float32, mode)
        S = pga32.subspace
        out.append((pga32, lambda r, p: r >> p, [(S.even_grade(), 'single'), (S.antivector(), 'batch')]))
        out.append((pga32, lambda a, b: a * b, [(S.even_grade(), 'batch'), (S.even_grade(), 'batch')]))
        out.append((pga32, lambda b: b.exp().normalized().motor_log(), [(S.bivector(), 'batch')]))
        cga = ctx_of((4, 1, 0), torch.float32, mode)
        F = cga.subspace.full()
        out.append((cga, lambda a, b: a * b, [(F, 'batch'), (F, 'batch')]))
    pga64 = ctx_of((3, 0, 1), torch.float64, 'fast')
    out.append((pga64, lambda b: b.exp().normalized().motor_log(), [(pga64.subspace.bivector(), 'batch')]))
    return out


def _ctx_factory():
    from numga_b200.backend import B200Context
    contexts = {}

    def ctx_of(pqr, dtype, mode):
        key = (pqr, dtype, mode)
        if key not in contexts:
            contexts[key] = B200Context(pqr, dtype=dtype, mode=mode, device='cpu', compile_only=True)
        return contexts[key]
    return ctx_of


def _run(ctx, fn, specs, batch=8):
    import torch
    mvs = []
    for sub, kind in specs:
        shape = (batch, len(sub)) if kind == 'batch' else (len(sub),)
        mvs.append(ctx.multivector(values=torch.zeros(shape, dtype=ctx.dtype), subspace=sub))
    fn(*mvs)              # eager: one kernel per method call
    ctx.jit(fn)(*mvs)     # fused: one kernel for the whole expression


def jobs():
    """(kind, argument) pairs of the product-side programs; every job is independent of the others."""
    out = [('bench', None), ('gather', None)]
    out += [('physics', (dt, mode)) for dt in ('float32', 'float64') for mode in ('fast', 'faithful')]
    return out


def run_job(job):
    _paths()
    import torch
    kind, arg = job
    ctx_of = _ctx_factory()
    if kind == 'bench':
        for ctx, fn, specs in bench_programs(ctx_of):
            _run(ctx, fn, specs)
    elif kind == 'gather':
        # gather / scatter through `indexed` operands and `out=` destinations (tests/test_gpu_parity.py)
        import numpy as np
        from numga_b200 import runtime as rt
        for dtype in (torch.float32, torch.float64):
            ctx = ctx_of((3, 0, 1), dtype, 'faithful')
            M = ctx.subspace.even_grade()
            a, b, o = (ctx.multivector(values=torch.zeros((8, 8), dtype=dtype), subspace=M) for _ in range(3))
            idx = torch.arange(8)
            ctx.jit(lambda x, y: x * y)(a.indexed(idx), b, out=o.indexed(idx))
            op = ctx.algebra.operator.product(M, M)
            terms = np.concatenate([op.index, op.coeff[:, None]], axis=1)
            rt.Program.from_terms(terms, [8, 8], 8, rt.F32 if dtype == torch.float32 else rt.F64, flags=rt.FAITHFUL)
    elif kind == 'physics':
        # rigid-body physics (config 5): the generic per-method kernels and the six fused kernels of a sub-step
        from numga_b200.examples.physics import ChainStep, FusedStep, setup_bodies
        ctx = ctx_of('x+y+z+w0', getattr(torch, arg[0]), arg[1])
        bodies, sets = setup_bodies(ctx, n_bodies=12)
        bodies.integrate(1 / 200, sets)
        FusedStep(bodies, sets).integrate(bodies, 1 / 200)
        for bpt in (None, 84):          # the chain-resident pipeline kernel: default tile and the partial-tile test's
            ChainStep(bodies, sets, bodies_per_tile=bpt).integrate(bodies, 1 / 200)
        bodies.kinetic_energy()
    return job


def rt_cache_dir():
    return os.environ.get('NGB200_CACHE_DIR') or os.path.join(os.path.dirname(os.path.abspath(__file__)), '_kcache')


def _dispatch(item):
    """Pool entry point: (module name | None, job).  Extra providers are modules with `jobs()` and `run_job(job)`."""
    module, job = item
    _paths()
    if module is None:
        return run_job(job)
    path, name = module
    if path not in sys.path:
        sys.path.insert(0, path)
    return importlib.import_module(name).run_job(job)


def prewarm(workers=None, extra=()):
    """Compile everything; `extra` = [(directory, module name)] of additional job providers.  Returns the number of
    cubins in the cache afterwards."""
    import multiprocessing as mp
    _paths()
    todo = [(None, j) for j in jobs()]
    for path, name in extra:
        if path not in sys.path:
            sys.path.insert(0, path)
        mod = importlib.import_module(name)
        todo += [((path, name), j) for j in mod.jobs()]
    todo.sort(key=lambda t: -getattr(importlib.import_module(t[0][1]), 'weight', lambda j: 0)(t[1]) if t[0] else -1000)
    workers = workers or max(1, min(len(todo), os.cpu_count() or 1, 16))
    if workers == 1:
        for t in todo:
            _dispatch(t)
    else:
        with mp.get_context('spawn').Pool(workers) as pool:
            for _ in pool.imap_unordered(_dispatch, todo, chunksize=1):
                pass
    d = rt_cache_dir()
    return len([f for f in os.listdir(d) if f.endswith('.cubin')]) if os.path.isdir(d) else 0


if __name__ == '__main__':
    print(prewarm(), 'kernels in cache')
