"""Compile (NVRTC, sm_100a) the generated kernels of the standard programs into the in-tree cubin cache.

Runs without a GPU: contexts are created in compile-only mode, the expressions of the case catalogue
(tests/cases.py, if present) and of bench.py are executed on uninitialised host tensors so that every
kernel they need — per-method eager kernels, fused `jit` kernels, operator-table kernels — is generated,
compiled and stored under numga_b200/_kcache/.  On the GPU box these are cache hits (no JIT there).
"""
import os
import sys

import numpy as np
import torch

from numga_b200 import runtime as rt
from numga_b200.backend import B200Context

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    tests = os.path.join(ROOT, 'tests')
    if not os.path.isdir(tests):
        return []
    if tests not in sys.path:
        sys.path.insert(0, tests)
    from cases import CASES
    return CASES


def bench_programs(ctx_of):
    """The fused programs bench.py times (kept in one place so build() and bench agree)."""
    out = []
    pga32 = ctx_of((3, 0, 1), torch.float32, 'fast')
    S = pga32.subspace
    out.append((pga32, lambda r, p: r >> p, [(S.even_grade(), 'single'), (S.antivector(), 'batch')]))
    out.append((pga32, lambda a, b: a * b, [(S.even_grade(), 'batch'), (S.even_grade(), 'batch')]))
    out.append((pga32, lambda b: b.exp().normalized().motor_log(), [(S.bivector(), 'batch')]))
    pga64 = ctx_of((3, 0, 1), torch.float64, 'fast')
    out.append((pga64, lambda b: b.exp().normalized().motor_log(), [(pga64.subspace.bivector(), 'batch')]))
    cga = ctx_of((4, 1, 0), torch.float32, 'fast')
    F = cga.subspace.full()
    out.append((cga, lambda a, b: a * b, [(F, 'batch'), (F, 'batch')]))
    return out


def prewarm(batch=8):
    contexts = {}

    def ctx_of(pqr, dtype, mode):
        key = (pqr, dtype, mode)
        if key not in contexts:
            contexts[key] = B200Context(pqr, dtype=dtype, mode=mode, device='cpu', compile_only=True)
        return contexts[key]

    before = len(os.listdir(rt_cache_dir())) if os.path.isdir(rt_cache_dir()) else 0

    def run(ctx, fn, specs):
        mvs = []
        for sub, kind in specs:
            shape = (batch, len(sub)) if kind == 'batch' else (len(sub),)
            mvs.append(ctx.multivector(values=torch.zeros(shape, dtype=ctx.dtype), subspace=sub))
        fn(*mvs)              # eager: one kernel per method call
        ctx.jit(fn)(*mvs)     # fused: one kernel for the whole expression

    for name, pqr, specs, scale, fn in _cases():
        for dtype in (torch.float32, torch.float64):
            for mode in ('fast', 'faithful'):
                ctx = ctx_of(pqr, dtype, mode)
                run(ctx, fn, [(getattr(ctx.subspace, s)(), 'batch' if k == 'batch' else 'single') for s, k in specs])
    for ctx, fn, specs in bench_programs(ctx_of):
        run(ctx, fn, specs)
    # rigid-body physics (config 5): the generic per-method kernels and the six fused kernels of a sub-step
    from numga_b200.examples.physics import FusedStep, setup_bodies
    for dtype in (torch.float32, torch.float64):
        for mode in ('fast', 'faithful'):
            ctx = ctx_of('x+y+z+w0', dtype, mode)
            bodies, sets = setup_bodies(ctx, n_bodies=12)
            bodies.integrate(1 / 200, sets)
            FusedStep(bodies, sets).integrate(bodies, 1 / 200)
            bodies.kinetic_energy()
    after = len(os.listdir(rt_cache_dir())) if os.path.isdir(rt_cache_dir()) else 0
    return max(after, before) // 2


def rt_cache_dir():
    return os.environ.get('NGB200_CACHE_DIR') or os.path.join(os.path.dirname(os.path.abspath(__file__)), '_kcache')


if __name__ == '__main__':
    print(prewarm(), 'kernels in cache')
