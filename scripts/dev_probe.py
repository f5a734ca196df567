"""Developer probe: raw C-ABI kernels on a GPU — correctness vs numpy einsum + bandwidth sweep.
Usage: python scripts/dev_probe.py [quick]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from numga_b200.algebra import Algebra
from numga_b200 import runtime as rt

def terms_of(op):
    return np.concatenate([op.index, op.coeff[:, None]], axis=1)

def soa(n, N, dtype=torch.float32, seed=0):
    g = torch.Generator(device='cuda'); g.manual_seed(seed)
    return torch.randn((n, N), device='cuda', dtype=dtype, generator=g)

def ops(t):
    return (t.data_ptr(), t.stride(0), t.stride(1) if t.dim() > 1 else 0)

def check(name, op, bcast_mask, dtype, N=1003):
    tdt = torch.float32 if dtype == rt.F32 else torch.float64
    n_in = [len(a) for a in op.inputs]; n_out = len(op.output)
    for flags in (rt.FAST, rt.FAITHFUL):
        prog = rt.Program.from_terms(terms_of(op), n_in, n_out, dtype, bcast_mask, flags)
        xs = [soa(n, 1 if (bcast_mask >> k) & 1 else N, tdt, seed=k) for k, n in enumerate(n_in)]
        y = torch.full((n_out, N), float('nan'), device='cuda', dtype=tdt)
        prog.launch([ops(x) for x in xs], [ops(y)], N, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        K = op.kernel
        letters = 'abcdefg'[:len(n_in)]
        expr = '...' + letters + 'z,' + ','.join('...' + l for l in letters) + '->...z'
        ref = np.einsum(expr, K, *[x.cpu().numpy().T for x in xs], optimize=True)
        got = y.cpu().numpy().T
        err = np.abs(got - ref).max() / np.abs(ref).max()
        exact = np.array_equal(got, ref.astype(got.dtype))
        print(f'  check {name} flags={flags} dtype={dtype} relerr={err:.3e} bitexact={exact} info={prog.info}', flush=True)

def bench(name, op, bcast_mask, dtype, N, iters=10, bytes_per_item=None):
    tdt = torch.float32 if dtype == rt.F32 else torch.float64
    es = 4 if dtype == rt.F32 else 8
    n_in = [len(a) for a in op.inputs]; n_out = len(op.output)
    prog = rt.Program.from_terms(terms_of(op), n_in, n_out, dtype, bcast_mask, rt.FAST)
    xs = [soa(n, 1 if (bcast_mask >> k) & 1 else N, tdt, seed=k) for k, n in enumerate(n_in)]
    y = torch.empty((n_out, N), device='cuda', dtype=tdt)
    stream = torch.cuda.current_stream().cuda_stream
    args = ([ops(x) for x in xs], [ops(y)], N, stream)
    for _ in range(3): prog.launch(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): prog.launch(*args)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bpi = es * (sum(n for k, n in enumerate(n_in) if not (bcast_mask >> k) & 1) + n_out)
    gbs = bpi * N / ms / 1e6
    print(f'  bench {name} N={N} {ms:.3f} ms  {N/ms/1e6:.2f} G items/s  {gbs:.0f} GB/s  regs={prog.info["regs_vec"]}/{prog.info["regs_scalar"]} vec={prog.info["vec"]}', flush=True)
    del xs, y
    return gbs

def main():
    quick = 'quick' in sys.argv
    print(rt.version(), torch.cuda.get_device_name(0), {k: v for k, v in os.environ.items() if k.startswith('NGB200')}, flush=True)
    A = Algebra.from_pqr(3, 0, 1); M = A.subspace.even_grade(); P = A.subspace.antivector()
    Cg = Algebra.from_pqr(4, 1, 0); F = Cg.subspace.full()
    gp, sw, cga = A.operator.product(M, M), A.operator.sandwich(M, P), Cg.operator.product(F, F)
    if 'check' in sys.argv or not quick:
        for dt in (rt.F32, rt.F64):
            check('gp(M,M)', gp, 0, dt)
            check('sandwich(M,P) bcast', sw, 0b101, dt)
            check('sandwich(M,P) batched', sw, 0, dt)
            check('cga gp(F,F)', cga, 0, dt)
    # copy roofline probe with torch for reference
    a = torch.empty(1 << 30, device='cuda', dtype=torch.float32); b = torch.empty_like(a)
    for _ in range(3): b.copy_(a)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    print(f'  torch copy 4GiB: {8 * (1 << 30) / (e0.elapsed_time(e1) / 10) / 1e6:.0f} GB/s', flush=True)
    del a, b
    res = {}
    res['gp'] = bench('gp(M,M) f32', gp, 0, rt.F32, 1 << 26)
    res['gp1M'] = bench('gp(M,M) f32', gp, 0, rt.F32, 1 << 20, iters=50)
    res['sw'] = bench('sandwich bcast f32', sw, 0b101, rt.F32, 1 << 28)
    res['cga'] = bench('cga gp(F,F) f32', cga, 0, rt.F32, 1 << 24)
    if not quick:
        res['gp64'] = bench('gp(M,M) f64', gp, 0, rt.F64, 1 << 25)
        res['swb'] = bench('sandwich batched f32', sw, 0, rt.F32, 1 << 26)
    print('RESULT', json.dumps({'env': {k: v for k, v in os.environ.items() if k.startswith('NGB200')}, **res}), flush=True)

main()
