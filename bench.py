#!/usr/bin/env python
"""Benchmark of the hot path (see BASELINE.json): batched multivector products on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs]

Headline workload (BASELINE.json configs[1]): 3D PGA (3,0,1), sandwich R >> p of 2^28 points with ONE
broadcast motor, fp32.  A step = one pass of the generated kernel over the batch.  One process per GPU
(torchrun for N > 1); the batch axis is sharded with no collective on the data path ("weak": every GPU
processes its own 2^28-point slice).  Prints ONE JSON line on rank 0.

  value          points/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e            same metric through the C-ABI host-buffer call (ngb200_program_run_host): pinned host inputs,
                 H2D + kernel + D2H inside the timed region
  roofline       algorithmic bytes (32 B/point) / kernel time vs the measured copy bandwidth
  cpu_baseline   the oracle (numpy einsum restatement of the reference's numpy backend) on the host cores
  configs        (with --configs) the other BASELINE.json configurations, device timed, each with its roofline

`--impl reference` times the reference's CPU algorithm (oracle port; the reference is a Python package that
cannot travel to the GPU box) on all host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POINTS_PER_GPU = 1 << 28
BYTES_PER_POINT = 32          # 4 components read + 4 written, fp32; the broadcast motor counts 0
METRIC = 'multivector products/sec (3D PGA motor sandwich R >> p, broadcast motor, fp32)'


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons polled through NVML from a background thread every ~2 ms, so that even a
    25 ms timed region gets samples DURING it (nvidia-smi's own loop cannot go below ~50 ms; it is the fallback when
    pynvml is unavailable, in which case the same kernel is kept running, untimed, to observe the clocks under load)."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')
    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t_begin = self.t_end = None
        self.nvml = None
        self.stop = False
        self.source = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.source = 'nvml, ~2 ms period'
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = 'nvidia-smi -lms 50'
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5:      # wait until the sampler is actually producing
                time.sleep(0.01)
        except Exception:
            self.proc = None
        return self

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mw = n.nvmlDeviceGetPowerUsage(self.handle)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                flags = ['Active' if mask & bit else 'Not Active' for _, bit in self.REASONS]
                self.rows.append((time.time(), [str(sm), str(self.max_sm), str(mw / 1000.0)] + flags))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def begin(self):
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    def samples_in_region(self):
        return [r for t, r in self.rows if self.t_begin is not None and t >= self.t_begin and (self.t_end is None or t <= self.t_end + (0.0 if self.nvml else 0.05))]

    def __exit__(self, *exc):
        self.stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, note=None):
        sm, mx, reasons, power = [], 0, set(), []
        for r in self.samples_in_region():
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1])); power.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        out = {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
               'samples': len(sm), 'power_w_max': max(power) if power else None, 'source': self.source}
        if note:
            out['note'] = note
        return out


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the oracle on the host cores
# ---------------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(n_points, seed):
    from oracle import numga_oracle as O
    alg = O.Algebra((3, 0, 1))
    rng = np.random.default_rng(seed)
    _W['R'] = O.MV.make(alg, 'even_grade', rng.standard_normal(8), np.float32)
    _W['p'] = O.MV.make(alg, 'antivector', rng.standard_normal((n_points, 4)), np.float32)
    _W['R'] >> _W['p']          # builds and caches the integer tables


def _cpu_worker_step(_):
    t = time.perf_counter()
    out = _W['R'] >> _W['p']
    return time.perf_counter() - t, float(out.values[0, 0])


def cpu_reference(steps, warmup, points_per_worker=1 << 22, max_workers=64):
    """Reference algorithm on the host: numpy einsum is single-threaded, so one process per core, each with
    its own slice (the same sharding the GPUs use). Returns (points/s, cores, sample description, ms/step)."""
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, max_workers))
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(points_per_worker, 1)) as pool:
        for _ in range(warmup):
            pool.map(_cpu_worker_step, range(cores))
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_worker_step, range(cores))
        dt = (time.perf_counter() - t0) / steps
    total = points_per_worker * cores
    return total / dt, cores, f'{cores} processes x 2^{int(np.log2(points_per_worker))} points per step (oracle port of the numpy backend, einsum optimize=True, order=F)', dt * 1e3


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    value, cores, sample, ms = cpu_reference(max(1, args.steps), max(1, args.warmup))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'points/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE.json configs[1]: 3D PGA (3,0,1) R >> p, broadcast motor, fp32 (bounded CPU sample per step)'},
        'cpu_baseline': {'value': value, 'unit': 'points/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------------
def device_mv(ctx, sub, n, seed, rank, scale=1.0):
    """Synthetic multivector batch generated ON the device: every rank fills its own slice of one
    counter-based normal stream (offset = rank * n), component rows are independent streams."""
    import torch
    from numga_b200 import runtime as rt
    from numga_b200.backend import soa_empty
    v = soa_empty(len(sub), (n,), ctx.dtype, ctx.device)
    rows = v.movedim(-1, 0)
    dt = rt.F32 if ctx.dtype == torch.float32 else rt.F64
    for c in range(len(sub)):
        rt.fill_normal(rows[c].data_ptr(), n, seed=seed * 1000 + c, offset=rank * n, scale=scale, dtype=dt,
                       stream=torch.cuda.current_stream().cuda_stream)
    return ctx.multivector(values=v, subspace=sub)


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPUs NVML reports as local to the GPU, so that the page-locked staging buffers of the
    host path are allocated (first touch) on the NUMA node the GPU's PCIe root hangs off.  Returns the CPU count or
    None when NVML / affinity is unavailable (the run is still valid, just not NUMA-local)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def time_steps(fn, steps, warmup, dist=None):
    """CUDA-event timing on the launching stream, barrier + synchronize on both sides, max over ranks."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps


def other_configs(torch, B200Context, rank, peak):
    """The remaining BASELINE.json configurations on one GPU, device timed (supplementary)."""
    out = []

    def entry(name, unit_bytes, n, ms, flops=None, note=None):
        gbs = unit_bytes * n / ms / 1e6
        e = {'config': name, 'items': n, 'ms': ms, 'items_per_s': n / ms * 1e3, 'bytes_per_item': unit_bytes,
             'achieved_gbs': gbs, 'frac_of_hbm_peak': gbs / peak}
        if flops:
            e['tflops'] = flops * n / ms / 1e9
        if note:
            e['note'] = note
        out.append(e)

    pga = B200Context((3, 0, 1), dtype=torch.float32)
    S = pga.subspace
    for n, tag in ((1 << 20, 'config 1: motor*motor, batch 2^20 (96 MB: L2-resident between launches)'),
                   (1 << 26, 'config 1: motor*motor, batch 2^26')):
        a, b = device_mv(pga, S.even_grade(), n, 11, rank), device_mv(pga, S.even_grade(), n, 12, rank)
        ms = time_steps(lambda: a * b, 20, 3)
        entry(tag, 96, n, ms, flops=88)
        if n == 1 << 20:
            # the eager call above is bound by the python/driver launch path (~35 us); 20 products recorded in one CUDA
            # graph show the kernel itself
            def twenty(a, b):
                for _ in range(20):
                    out = a * b
                return out
            g = pga.capture(twenty, a, b)
            ms = time_steps(g.replay, 10, 3) / 20
            entry('config 1: motor*motor, batch 2^20, 20 launches per CUDA graph replay (B200Context.capture)', 96, n, ms, flops=88)
        del a, b
    for dtype, tag, ub in ((torch.float32, 'fp32', 48), (torch.float64, 'fp64', 96)):
        c = B200Context((3, 0, 1), dtype=dtype)
        n = 1 << 28 if dtype == torch.float32 else 1 << 27
        bv = device_mv(c, c.subspace.bivector(), n, 13, rank, scale=0.5)
        f = c.jit(lambda x: x.exp().normalized().motor_log())
        ms = time_steps(lambda: f(bv), 5, 2)
        prog = c.engine.cache[next(k for k in c.engine.cache if k[0][0] == 'jit')].program
        entry(f'config 3: exp->normalized->motor_log fused round trip, {tag}, batch 2^{int(np.log2(n))}', ub, n, ms,
              flops=prog.info['n_instr'], note='compute-bound: one kernel, all intermediates in registers; flops = live IR instructions')
        del bv
    # exp alone: the generic bisection algorithm (the reference's default dispatch) vs the opt-in closed form
    # (numga_b200.optimized, the counterpart of the reference's optional extension/optimized.py)
    import numga_b200.optimized as opt
    c = B200Context((3, 0, 1), dtype=torch.float32)
    n = 1 << 28
    bv = device_mv(c, c.subspace.bivector(), n, 13, rank, scale=0.5)
    for tag, on in (('generic bisection (default dispatch)', False), ('opt-in closed form (numga_b200.optimized)', True)):
        (opt.enable if on else opt.disable)()
        f = c.jit(lambda x: x.exp())
        ms = time_steps(lambda: f(bv), 5, 2)
        prog = c.engine.cache[next(k for k in c.engine.cache if k[0][0] == 'jit')].program
        entry(f'config 3 part: bivector exp only, fp32, batch 2^28, {tag}', 56, n, ms, flops=prog.info['n_instr'])
    opt.disable()
    del bv
    cga = B200Context((4, 1, 0), dtype=torch.float32)
    F = cga.subspace.full()
    n = 1 << 26
    x, y = device_mv(cga, F, n, 14, rank), device_mv(cga, F, n, 15, rank)
    ms = time_steps(lambda: x * y, 10, 3)
    entry('config 4: CGA (4,1,0) full*full geometric product, batch 2^26 (SIMT, 1024 FMA/item)', 384, n, ms, flops=2016)
    # the tensor-core formulation of the same product, for the comparison BASELINE.json configs[3] asks for:
    # materialise the outer product P[b, (i,j)] = a_i b_j and multiply by the [1024, 32] sign matrix with cuBLAS
    # (TF32 tensor cores).  32x the flops of the sparse form and 4 KB of intermediate per item; library call, NOT the
    # product path.
    nb = 1 << 20
    K = torch.zeros(1024, 32, device='cuda')
    op = cga.algebra.operator.product(F, F)
    idx = torch.as_tensor(op.index.astype(np.int64), device='cuda')
    K[idx[:, 0] * 32 + idx[:, 1], idx[:, 2]] = torch.as_tensor(op.coeff.astype(np.float32), device='cuda')
    xa, ya = x.values[:nb].contiguous(), y.values[:nb].contiguous()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True

    def tensor_core_form():
        return (xa[:, :, None] * ya[:, None, :]).reshape(nb, 1024) @ K
    ms_tc = time_steps(tensor_core_form, 5, 2)
    ref = cga.multivector(values=xa, subspace=F) * cga.multivector(values=ya, subspace=F)
    err_tc = float(((tensor_core_form() - ref.values).abs().amax(dim=-1) / ref.values.abs().amax(dim=-1)).max().item())
    torch.backends.cuda.matmul.allow_tf32 = prev
    e = {'config': 'config 4 comparison: tensor-core formulation (outer product + cuBLAS TF32 GEMM with the [1024,32] sign matrix), batch 2^20',
         'items': nb, 'ms': ms_tc, 'items_per_s': nb / ms_tc * 1e3, 'bytes_per_item': 384, 'achieved_gbs': 384 * nb / ms_tc / 1e6,
         'frac_of_hbm_peak': 384 * nb / ms_tc / 1e6 / peak, 'tflops': 2 * 1024 * 32 * nb / ms_tc / 1e9,
         'max_rel_error_vs_simt': err_tc, 'simt_items_per_s': out[-1]['items_per_s'],
         'note': 'every K[i,:,:] is a signed permutation: there is no contraction to amortise, so the GEMM form does 32x the flops '
                 'and moves 4 KB/item of intermediate; TF32 also misses the 1e-5 parity gate. SIMT wins; tensor cores are not used.'}
    out.append(e)
    del x, y, xa, ya
    for dtype, tag in ((torch.float32, 'fp32'), (torch.float64, 'fp64')):
        out.append(physics_config(torch, B200Context, rank, peak, dtype, tag))
    out.append(training_step_config(torch, B200Context, rank, peak))
    out.extend(generic_physics_configs(torch, B200Context, rank, peak))
    return out


def generic_physics_configs(torch, B200Context, rank, peak, n_target=1 << 22):
    """Supplementary: the rigid-body step written exactly like the reference (Body.integrate, one multivector method
    after the other), run eagerly (one kernel per method) and with automatic fusion (B200Context(lazy=True))."""
    from numga_b200 import runtime as rt
    from numga_b200.examples.physics import replicate_chains, setup_bodies
    out = []
    n_chains = n_target // 12
    for mode in ('eager', 'lazy'):
        ctx = B200Context('x+y+z+w0', dtype=torch.float32, lazy=(mode == 'lazy'))
        bodies, sets = setup_bodies(ctx, n_bodies=12)
        big, bsets = replicate_chains(bodies, sets, n_chains)
        big = big.copy(rate=big.rate + device_mv(ctx, ctx.subspace.bivector(), 12 * n_chains, 16, rank, scale=0.3))
        big.rate.values
        state = [big]

        def step():
            new = state[0].integrate(1 / 200, bsets)
            new.motor.values, new.rate.values
            state[0] = new
        l0 = rt.launch_count()
        step()
        launches = rt.launch_count() - l0
        ms = time_steps(step, 5, 2)
        n = 12 * n_chains
        out.append({'config': f'supplementary: reference-style Body.integrate (unchanged formulas, {mode}: {launches} generated kernels per sub-step), fp32, {n} bodies',
                    'items': n, 'ms': ms, 'items_per_s': n / ms * 1e3, 'bytes_per_item': 0, 'achieved_gbs': 0.0, 'frac_of_hbm_peak': 0.0,
                    'kernel_launches_per_step': launches})
        del big, bsets, state
        torch.cuda.empty_cache()
    return out


def training_step_config(torch, B200Context, rank, peak, n=1 << 26):
    """Supplementary: forward + backward of the headline operator inside torch.autograd (fit ONE motor to n point
    correspondences).  The backward pass is the generated VJP kernel; the motor's gradient is reduced in-kernel."""
    ctx = B200Context((3, 0, 1), dtype=torch.float32)
    S = ctx.subspace
    p = device_mv(ctx, S.antivector(), n, 17, rank)
    q = ctx.multivector.bivector(np.array([0.3, -0.2, 0.1, 0.5, 0.2, -0.4], np.float32)).exp() >> p
    b = torch.zeros(6, device='cuda', requires_grad=True)

    def sq_err(r, x, y):
        d = ((r >> x) - y).values
        return ctx.multivector.scalar(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2] + d[..., 3] * d[..., 3])
    err = ctx.jit(sq_err)

    def step():
        b.grad = None
        loss = err(ctx.multivector(values=b, subspace=S.bivector()).exp(), p, q).values.sum()
        loss.backward()
    ms = time_steps(step, 10, 3)
    nbytes = (16 + 16 + 4 + 4 + 16 + 16 + 4) * n          # fwd: x, y in, err out; torch sum; bwd: x, y, cotangent in
    return {'config': f'supplementary: autograd training step (fit one motor to 2^{int(np.log2(n))} points: forward + torch sum + generated VJP kernel with in-kernel reduction)',
            'items': n, 'ms': ms, 'items_per_s': n / ms * 1e3, 'bytes_per_item': nbytes / n, 'achieved_gbs': nbytes / ms / 1e6,
            'frac_of_hbm_peak': nbytes / ms / 1e6 / peak, 'grad_finite': bool(torch.isfinite(b.grad).all().item())}


def physics_config(torch, B200Context, rank, peak, dtype, tag, n_target=1 << 24):
    """BASELINE.json configs[4]: the rigid-body core step (numga/examples/physics/core.py Body.integrate with two
    constraint colour sets) on ~2^24 bodies = independent 12-body chains, as SIX fused kernels per sub-step."""
    from numga_b200 import runtime as rt
    from numga_b200.examples.physics import FusedStep, replicate_chains, setup_bodies
    c = B200Context('x+y+z+w0', dtype=dtype)
    bodies, sets = setup_bodies(c, n_bodies=12)
    n_chains = n_target // 12
    big, bsets = replicate_chains(bodies, sets, n_chains)
    n_bodies = 12 * n_chains
    n_constraints = sum(int(s.body_idx.shape[1]) for s in bsets)
    big = big.copy(rate=big.rate + device_mv(c, c.subspace.bivector(), n_bodies, 16, rank, scale=0.3))
    fs = FusedStep(big, bsets)
    state = [big]

    def step():
        state[0] = fs.integrate(state[0], 1 / 200)
    l0 = rt.launch_count()
    step()
    launches = rt.launch_count() - l0
    ms = time_steps(step, 10, 3)
    finite = bool(torch.isfinite(state[0].motor.values).all().item() and torch.isfinite(state[0].rate.values).all().item())
    # compulsory traffic of the six-kernel split: live input components + outputs (+ gather indices) per item
    es = 4 if dtype == torch.float32 else 8
    words, instr = {}, {}
    for k, v in c.engine.cache.items():
        if k[0][0] != 'jit':
            continue
        name = k[0][1].__name__
        live_in = sum(1 for op, a, b, _ in v.ir['instr'] if op == rt.OP_INPUT and v.program.in_mode[a] != rt.BROADCAST)
        words[name] = live_in + sum(v.ir['out_width'])
        instr[name] = v.program.info['n_instr']
    per_body = (words['pre'] + words['post']) * es
    per_constraint = (words['apply'] + words['v_apply']) * es + 2 * 2 * 8      # two int64 body indices, two kernels
    total = per_body * n_bodies + per_constraint * n_constraints
    flops = (instr['pre'] + instr['post']) * n_bodies + (instr['apply'] + instr['v_apply']) * n_constraints
    gbs = total / ms / 1e6
    return {'config': f'config 5: rigid-body XPBD sub-step (Body.integrate, 2 colour sets), {tag}, {n_bodies} bodies in {n_chains} chains, '
                      f'{launches} fused kernels per sub-step', 'items': n_bodies, 'ms': ms, 'items_per_s': n_bodies / ms * 1e3,
            'bytes_per_item': total / n_bodies, 'achieved_gbs': gbs, 'frac_of_hbm_peak': gbs / peak,
            'tflops': flops / ms / 1e9, 'kernel_launches_per_step': launches, 'state_finite': finite,
            'note': 'bytes = live inputs + outputs of the six kernels (pre, apply x2, post, v_apply x2) incl. gathered inertia '
                    'matrices; the state-once minimum of SURVEY.md 8d (728 B fp32) would need a single pass'}


def run_ours(args):
    import torch
    from numga_b200 import B200Context, runtime as rt

    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (numga_b200 has no CPU path); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints a version banner on stdout when the communicator is created; stdout must carry ONE JSON line, so
        # file descriptor 1 points at stderr while the process group and its first collective come up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    peak, peak_src = measured_peaks()
    n = args.points
    ctx = B200Context((3, 0, 1), dtype=torch.float32, device=f'cuda:{local}')
    S = ctx.subspace
    R = ctx.multivector.motor(np.random.default_rng(1).standard_normal(8).astype(np.float32)).normalized()
    p = device_mv(ctx, S.antivector(), n, seed=2, rank=rank)
    sandwich = ctx.jit(lambda r, x: r >> x)
    sandwich(R, p)
    torch.cuda.synchronize()

    with ClockSampler(local) as clocks:
        for _ in range(args.warmup):
            sandwich(R, p)
        launches0 = rt.launch_count()
        clocks.begin()
        ms = time_steps(lambda: sandwich(R, p), args.steps, 0, dist)
        launches = rt.launch_count() - launches0                    # launches inside the timed region
        clocks.end()
        note = None
        if len(clocks.samples_in_region()) < 4:
            # the timed region (steps x ~1.3 ms) is shorter than nvidia-smi's sampling period: keep the identical
            # kernel running, untimed, so the clocks under this load are observed
            t0 = time.time()
            while time.time() - t0 < 1.0:
                for _ in range(20):
                    sandwich(R, p)
                torch.cuda.synchronize()
            note = 'timed region shorter than the sampling period; clocks sampled over the timed region plus a 1 s untimed continuation of the same kernel'
            clocks.end()
        clock_summary = clocks.summary(note)
    value = world * n / ms * 1e3
    achieved = BYTES_PER_POINT * n / ms / 1e6

    # ---- end to end through the C-ABI host-buffer call ---------------------------------------------------------
    import ctypes
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)      # pinned host buffers are first touched on the GPU's own NUMA node
    prog = ctx.engine.cache[next(k for k in ctx.engine.cache if k[0][0] == 'jit')].program
    e2e_n = n
    try:
        import psutil
        if psutil.virtual_memory().available < 6 * 2 * 16 * n * max(1, min(world, 8)):
            e2e_n = n // 8
    except Exception:
        pass
    hin, hout, hR = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    rt.check(rt.lib.ngb200_host_alloc(ctypes.byref(hin), 16 * e2e_n))
    rt.check(rt.lib.ngb200_host_alloc(ctypes.byref(hout), 16 * e2e_n))
    rt.check(rt.lib.ngb200_host_alloc(ctypes.byref(hR), 32))
    host_p = np.ctypeslib.as_array(ctypes.cast(hin, ctypes.POINTER(ctypes.c_float)), shape=(4, e2e_n))
    host_q = np.ctypeslib.as_array(ctypes.cast(hout, ctypes.POINTER(ctypes.c_float)), shape=(4, e2e_n))
    host_R = np.ctypeslib.as_array(ctypes.cast(hR, ctypes.POINTER(ctypes.c_float)), shape=(8,))
    host_R[:] = R.numpy()
    chunk = 1 << 22
    for i in range(0, e2e_n, chunk):        # pinned host input: a copy of the device data (SoA, as numpy order='F')
        host_p[:, i:i + chunk] = p.values[i:i + chunk].T.cpu().numpy()
    ins = [(hR.value, 1, 0), (hin.value, e2e_n, 1)]
    outs = [(hout.value, e2e_n, 1)]
    e2e_steps = max(1, min(args.steps, 10))
    prog.run_host(ins, outs, e2e_n, device=local)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prog.run_host(ins, outs, e2e_n, device=local)           # synchronous: returns when outputs are on the host
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    check = sandwich(R, p).values[:1 << 16].T.cpu().numpy()
    e2e_ok = bool(np.array_equal(check, host_q[:, :1 << 16]))
    e2e = {'value': world * e2e_n / e2e_s, 'unit': 'points/s', 'h2d_bytes_per_step': 16 * e2e_n + 32,
           'd2h_bytes_per_step': 16 * e2e_n, 'steps': e2e_steps, 'points_per_gpu': e2e_n,
           'matches_device_result': e2e_ok, 'numa_local_cpus': numa,
           'path': 'ngb200_program_run_host (pinned host SoA buffers, 128 MB chunks, H2D/kernel/D2H pipelined over 4 streams)'}
    for h in (hin, hout, hR):
        rt.lib.ngb200_host_free(h)
    os.sched_setaffinity(0, all_cpus)        # the CPU baseline below gets every host core back

    # ---- optional final gather over NVLink (not on the hot path; timed separately) ------------------------------
    gather = None
    if dist is not None:
        gn = 1 << 24
        src = sandwich(R, p).values[:gn].movedim(-1, 0).contiguous()
        dst = torch.empty((world,) + tuple(src.shape), device='cuda', dtype=src.dtype)
        g_ms = time_steps(lambda: dist.all_gather_into_tensor(dst, src), 5, 2, dist)
        gather = {'bytes_per_rank': src.numel() * 4, 'ms': g_ms, 'algbw_gbs': src.numel() * 4 * (world - 1) / g_ms / 1e6,
                  'note': 'NCCL all_gather of a 2^24-point output slice per rank; results stay sharded by default'}

    line = None
    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'r01_sandwich_traffic.json')
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get('dram_bytes_per_launch')
        line = {
            'metric': METRIC, 'value': value, 'unit': 'points/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': 'BASELINE.json configs[1]: 3D PGA (3,0,1) motor sandwich R >> p, one broadcast motor, fp32',
                       'points_per_gpu': n, 'layout': 'structure-of-arrays [4, n] in HBM', 'parallelism': f'batch-sharded x{world}, no collective',
                       'l2': f'inputs+outputs {2 * 16 * n / 1e9:.1f} GB per GPU per step >> 126 MB L2 (no flush needed)'},
            'clocks': clock_summary, 'gpu_launches': launches, 'e2e': e2e,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': BYTES_PER_POINT * n, 'kernel_ms': ms},
        }
        if gather:
            line['gather'] = gather
    if args.configs and world == 1:
        line['configs'] = other_configs(torch, B200Context, rank, peak)
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sample, _ = cpu_reference(3, 1)
        line['cpu_baseline'] = {'value': v, 'unit': 'points/s', 'cores': cores, 'kind': 'port', 'sample': sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--points', type=int, default=POINTS_PER_GPU, help='points per GPU (default 2^28, the BASELINE.json size)')
    ap.add_argument('--configs', action='store_true', help='also time the other BASELINE.json configurations (1 GPU)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
