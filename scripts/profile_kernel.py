"""Run ONE hot-path kernel a few times at benchmark size (target for ncu). Usage: profile_kernel.py <which> [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numga_b200 import B200Context
from bench import device_mv

which = sys.argv[1] if len(sys.argv) > 1 else 'sandwich'
launches = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 4
if which == 'sandwich':
    ctx = B200Context((3, 0, 1)); S = ctx.subspace
    R = ctx.multivector.motor(np.random.default_rng(1).standard_normal(8).astype(np.float32))
    p = device_mv(ctx, S.antivector(), 1 << 28, 2, 0)
    f = ctx.jit(lambda r, x: r >> x); run = lambda: f(R, p)
elif which == 'gp':
    ctx = B200Context((3, 0, 1)); S = ctx.subspace
    a, b = device_mv(ctx, S.even_grade(), 1 << 26, 3, 0), device_mv(ctx, S.even_grade(), 1 << 26, 4, 0)
    run = lambda: a * b
elif which == 'cga':
    ctx = B200Context((4, 1, 0)); F = ctx.subspace.full()
    a, b = device_mv(ctx, F, 1 << 24, 5, 0), device_mv(ctx, F, 1 << 24, 6, 0)
    run = lambda: a * b
elif which in ('roundtrip_closed', 'roundtrip_closed64'):
    import numga_b200.optimized as opt
    opt.enable()                      # opt-in closed forms: exp / normalized (reference's optimized.py) + log_motor
    ctx = B200Context((3, 0, 1), dtype=torch.float64 if which.endswith('64') else torch.float32)
    bv = device_mv(ctx, ctx.subspace.bivector(), 1 << 26, 7, 0, scale=0.5)
    f = ctx.jit(lambda x: x.exp().normalized().motor_log()); run = lambda: f(bv)
elif which in ('roundtrip', 'roundtrip64'):
    ctx = B200Context((3, 0, 1), dtype=torch.float64 if which.endswith('64') else torch.float32)
    bv = device_mv(ctx, ctx.subspace.bivector(), 1 << 24, 7, 0, scale=0.5)
    f = ctx.jit(lambda x: x.exp().normalized().motor_log()); run = lambda: f(bv)
if which == 'physics_time':
    import json
    from bench import physics_config, measured_peaks
    r = physics_config(torch, B200Context, 0, measured_peaks()[0], torch.float64 if 'f64' in sys.argv else torch.float32, 'fp')
    print(which, {k: v for k, v in os.environ.items() if k.startswith('NGB200') and 'CACHE' not in k}, f"{r['ms']:.3f} ms {r['items_per_s']/1e9:.3f} G/s {r['achieved_gbs']:.0f} GB/s")
    sys.exit(0)
if which == 'physics_generic_time':
    # the reference-style Body.integrate (numga/examples/physics/core.py, unchanged formulas) at benchmark size:
    # eager (one kernel per multivector method), lazy (automatic fusion), and the hand-arranged FusedStep
    from bench import time_steps
    from numga_b200 import runtime as rt
    from numga_b200.examples.physics import FusedStep, replicate_chains, setup_bodies
    n_chains = (1 << int(os.environ.get('LOGN', '22'))) // 12
    for mode in os.environ.get('MODES', 'eager,lazy,fused').split(','):
        ctx = B200Context('x+y+z+w0', lazy=(mode == 'lazy'))
        bodies, sets = setup_bodies(ctx, n_bodies=12)
        big, bsets = replicate_chains(bodies, sets, n_chains)
        big = big.copy(rate=big.rate + device_mv(ctx, ctx.subspace.bivector(), 12 * n_chains, 16, 0, scale=0.3))
        big.rate.values
        state = [big]
        if mode == 'fused':
            fs = FusedStep(big, bsets)
            def step(): state[0] = fs.integrate(state[0], 1 / 200)
        else:
            def step():
                new = state[0].integrate(1 / 200, bsets)
                new.motor.values; new.rate.values
                state[0] = new
        l0 = rt.launch_count(); step(); torch.cuda.synchronize(); launches = rt.launch_count() - l0
        ms = time_steps(step, 5, 2)
        print(which, mode, f'{12 * n_chains} bodies: {ms:.3f} ms per sub-step, {launches} generated-kernel launches, {12 * n_chains / ms / 1e6:.3f} G body-steps/s', flush=True)
        del big, bsets, state
        torch.cuda.empty_cache()
    sys.exit(0)
if which == 'cga_pitch':
    from bench import time_steps
    ctx = B200Context((4, 1, 0)); F = ctx.subspace.full(); n = 1 << 26
    from numga_b200 import runtime as rt
    for pad in (0, 32, 96, 256 + 32, 1024 + 32, 4096 + 32, 65536 + 32, 3 * 65536 + 96):
        def mk(seed):
            t = torch.empty((32, n + pad), dtype=torch.float32, device='cuda')
            for c in range(32):
                rt.fill_normal(t[c].data_ptr(), n, seed=seed * 1000 + c, stream=torch.cuda.current_stream().cuda_stream)
            return t
        ta, tb, to = mk(5), mk(6), torch.empty((32, n + pad), dtype=torch.float32, device='cuda')
        a = ctx.multivector(values=ta[:, :n].movedim(0, -1), subspace=F); b = ctx.multivector(values=tb[:, :n].movedim(0, -1), subspace=F)
        f = ctx.jit(lambda x, y: x * y)
        o = ctx.multivector(values=to[:, :n].movedim(0, -1), subspace=F)
        ms = time_steps(lambda: f(a, b, out=(o,)), 10, 3)
        print(which, 'pad', pad, f'{ms:.3f} ms {384*n/ms/1e6:.0f} GB/s', flush=True)
        del ta, tb, to, a, b, o
    sys.exit(0)
if which in ('cga_time', 'roundtrip_time'):
    import time
    from bench import time_steps
    if which == 'cga_time':
        ctx = B200Context((4, 1, 0)); F = ctx.subspace.full(); n = 1 << int(os.environ.get('LOGN', '26'))
        a, b = device_mv(ctx, F, n, 5, 0), device_mv(ctx, F, n, 6, 0)
        run = lambda: a * b; bpi = 384
    else:
        ctx = B200Context((3, 0, 1)); n = 1 << 27
        bv = device_mv(ctx, ctx.subspace.bivector(), n, 7, 0, scale=0.5)
        f = ctx.jit(lambda x: x.exp().normalized().motor_log()); run = lambda: f(bv); bpi = 48
    ms = time_steps(run, int(os.environ.get('STEPS', '10')), 3)
    prog = list(ctx.engine.cache.values())[-1].program
    print(which, {k: v for k, v in os.environ.items() if k.startswith('NGB200') and 'CACHE' not in k}, f'{ms:.3f} ms {n/ms/1e6:.2f} G/s {bpi*n/ms/1e6:.0f} GB/s', prog.info['regs_vec'], prog.info['regs_scalar'], prog.info['vec'])
    sys.exit(0)
for _ in range(launches):
    out = run()
torch.cuda.synchronize()
print('ok', which, launches, tuple(out.values.shape))
