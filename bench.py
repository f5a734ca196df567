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
# CPU side: the oracle (port of the reference's numpy backend) on the host cores
# ---------------------------------------------------------------------------------------------------------------
# Per configuration: items per worker per step for the reference's two numpy operator classes
# (NumpyEinsumOperator = the default, NumpySparseOperator; numga/backend/numpy/operator.py:102-160).  Sized so that
# one step is ~0.05-0.8 s on one core; the config sizes themselves (2^20 - 2^28) would take minutes to hours, numpy
# is single-threaded and the work is linear in the batch, so throughput is reported per item.
CPU_SAMPLES = {
    'cfg1': {'einsum': 1 << 18, 'sparse': 1 << 18},
    'cfg2': {'einsum': 1 << 22, 'sparse': 1 << 17},
    'cfg3_f32': {'einsum': 1 << 14, 'sparse': 1 << 14},
    'cfg3_f64': {'einsum': 1 << 14, 'sparse': 1 << 14},
    'cfg4': {'einsum': 1 << 12, 'sparse': 1 << 13},
    'cfg5_f32': {'einsum': 12 * 256, 'sparse': 12 * 256},
    'cfg5_f64': {'einsum': 12 * 256, 'sparse': 12 * 256},
}
PHYSICS_DT = 1 / 200


def physics_arrays(tag):
    """Set-up state of the 12-body chain exactly as the reference's `setup_bodies` produces it (committed fixture,
    generated from the unmodified reference by tests/golden/make_golden_physics.py)."""
    G = np.load(os.path.join(ROOT, 'tests', 'golden', 'physics.npz'))
    return {k[len(tag) + 1:]: G[k] for k in G.files if k.startswith(tag + '/')}


def cpu_workload(cfg, otype, n, seed):
    """One pass of configuration `cfg`'s hot path over n items with the oracle; returns (step, inputs, result)."""
    from oracle import numga_oracle as O
    O.OTYPE = otype
    rng = np.random.default_rng(seed)
    if cfg == 'cfg1':
        c = O.context((3, 0, 1), np.float32)
        a, b = (c.multivector('even_grade', rng.standard_normal((n, 8))) for _ in range(2))
        return lambda: a * b
    if cfg == 'cfg2':
        c = O.context((3, 0, 1), np.float32)
        R = c.multivector('even_grade', rng.standard_normal(8))
        p = c.multivector('antivector', rng.standard_normal((n, 4)))
        return lambda: R >> p
    if cfg.startswith('cfg3'):
        c = O.context((3, 0, 1), np.float32 if cfg.endswith('f32') else np.float64)
        bv = c.multivector('bivector', 0.5 * rng.standard_normal((n, 6)))
        return lambda: bv.exp().normalized().motor_log()
    if cfg == 'cfg4':
        c = O.context((4, 1, 0), np.float32)
        x, y = (c.multivector('multivector', rng.standard_normal((n, 32))) for _ in range(2))
        return lambda: x * y
    if cfg.startswith('cfg5'):
        from oracle import physics_oracle as P
        tag = cfg[-3:]
        arrays = physics_arrays(tag)
        chains = max(1, n // 12)
        rate = np.tile(arrays['init/rate'], (chains, 1)) + 0.3 * rng.standard_normal((12 * chains, 6))
        bodies, sets = P.from_arrays(arrays, np.float32 if tag == 'f32' else np.float64, n_chains=chains, rate=rate)
        return lambda: bodies.integrate(PHYSICS_DT, sets)
    raise KeyError(cfg)


def _farm_worker(conn, cpu):
    try:
        os.sched_setaffinity(0, {cpu})
    except Exception:
        pass
    sys.path.insert(0, ROOT)
    cache = {}
    while True:
        msg = conn.recv()
        if msg is None:
            return
        if msg[0] == 'prepare':
            _, cfg, otype, n, warmup = msg
            key = (cfg, otype, n)
            if key not in cache:
                cache.clear()
                cache[key] = cpu_workload(cfg, otype, n, seed=1000 + cpu)
            with np.errstate(all='ignore'):
                for _ in range(warmup):
                    cache[key]()
            conn.send('ready')
        else:
            _, cfg, otype, n, steps = msg
            step, times = cache[(cfg, otype, n)], []
            with np.errstate(all='ignore'):
                for _ in range(steps):
                    t0 = time.perf_counter()
                    step()
                    times.append(time.perf_counter() - t0)
            conn.send(times)


class CpuFarm:
    """One worker process per host core, each PINNED to its own CPU (numpy's einsum is single-threaded, so the batch
    is sharded over processes exactly like it is over GPUs).  A measurement = every participating worker prepares
    (builds inputs and integer tables, warm-up passes), then all start together and run `steps` passes; the
    throughput of step s is sum over workers of items / time, and the MEDIAN over steps is reported."""

    def __init__(self, max_workers=64):
        import multiprocessing as mp
        for v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
            os.environ[v] = '1'                      # one worker = one core, also inside BLAS-backed einsum paths
        cpus = sorted(os.sched_getaffinity(0))[:max_workers]
        ctx = mp.get_context('spawn')
        self.workers = []
        for cpu in cpus:
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_farm_worker, args=(child, cpu), daemon=True)
            p.start()
            self.workers.append((p, parent))
        self.cores = len(cpus)

    def measure(self, cfg, otype, n, workers, steps=3, warmup=1):
        use = self.workers[:workers]
        for _, c in use:
            c.send(('prepare', cfg, otype, n, warmup))
        for _, c in use:
            assert c.recv() == 'ready'
        for _, c in use:
            c.send(('go', cfg, otype, n, steps))
        times = np.array([c.recv() for _, c in use])          # [workers, steps]
        per_step = (n / times).sum(axis=0)
        return float(np.median(per_step)), float(np.median(times.max(axis=0)) * 1e3)

    def baselines(self, cfg, unit):
        """einsum + sparse operator, one core and all cores, for one configuration."""
        out = {'unit': unit, 'kind': 'port', 'cores': self.cores,
               'sample': {k: f'{v} items per worker per step' for k, v in CPU_SAMPLES[cfg].items()},
               'how': 'oracle port of the reference numpy backend; workers pinned one per CPU; median over 3 steps after 1 warm-up'}
        for otype in ('einsum', 'sparse'):
            n = CPU_SAMPLES[cfg][otype]
            out[f'{otype}_1core'], _ = self.measure(cfg, otype, n, 1)
            out[f'{otype}_allcores'], _ = self.measure(cfg, otype, n, self.cores)
        out['value'] = out['einsum_allcores']           # NumpyEinsumOperator is the reference's default otype
        return out

    def close(self):
        for p, c in self.workers:
            try:
                c.send(None)
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=5)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    farm = CpuFarm()
    n = CPU_SAMPLES['cfg2']['einsum']
    value, ms = farm.measure('cfg2', 'einsum', n, farm.cores, steps=max(1, args.steps), warmup=max(1, args.warmup))
    cores = farm.cores
    farm.close()
    sample = (f'{cores} pinned processes x 2^{int(np.log2(n))} points per step (oracle port of the numpy backend, '
              'NumpyEinsumOperator: einsum optimize=True, order=F); median step')
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


def rel_err(got, want):
    """max over items of ||got - want||inf / ||want||inf (SURVEY.md 8d parity gate)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = np.maximum(np.abs(want).max(axis=-1), 1e-300)
    with np.errstate(all='ignore'):
        e = np.abs(got - want).max(axis=-1) / scale
    return float(np.nanmax(e)) if e.size else 0.0


PARITY_N = 1 << 16
# value gates of the benchmarked (fast) mode against the reference at the same precision; derivation in
# profiles/r02_tolerances.md (config 5: (motor, rate) after ONE sub-step; scaled per sub-step count below)
TOL = {'cfg1': 1e-5, 'cfg2': 1e-5, 'cfg4': 1e-5, 'cfg3_f32': 2e-5, 'cfg3_f64': 2e-11,
       'cfg5_f32': (1e-5, 0.8e-3), 'cfg5_f64': (1e-12, 1e-11)}


def entry(cfg, name, n, ms, bytes_per_item, peak, unit='products/s', **extra):
    gbs = bytes_per_item * n / ms / 1e6
    e = {'id': cfg, 'config': name, 'items': n, 'ms': ms, 'items_per_s': n / ms * 1e3, 'unit': unit,
         'roofline': {'bound': 'hbm', 'bytes_per_item': bytes_per_item, 'achieved': gbs, 'peak': peak, 'unit': 'GB/s',
                      'frac': gbs / peak}}
    e.update(extra)
    return e


def run_cfg1(torch, B200Context, rank, peak, faithful=True):
    """configs[0]: 3D PGA motor * motor, batch 2^20 (BASELINE size; 96 MB working set: L2-resident) and 2^26."""
    from oracle import numga_oracle as O
    out = []
    pga = B200Context((3, 0, 1), dtype=torch.float32)
    S = pga.subspace
    rng = np.random.default_rng(101)
    ha, hb = rng.standard_normal((PARITY_N, 8)).astype(np.float32), rng.standard_normal((PARITY_N, 8)).astype(np.float32)
    oc = O.context((3, 0, 1), np.float32)
    want = (oc.multivector('even_grade', ha) * oc.multivector('even_grade', hb)).values
    err = rel_err((pga.multivector.motor(ha) * pga.multivector.motor(hb)).numpy(), want)
    for n in (1 << 20, 1 << 26):
        a, b = device_mv(pga, S.even_grade(), n, 11, rank), device_mv(pga, S.even_grade(), n, 12, rank)
        ms = time_steps(lambda: a * b, 50 if n == 1 << 20 else 20, 5)
        extra = {}
        if n == 1 << 20:
            # host cost of one eager call (python dispatch + output allocation + C launch), GPU idle
            torch.cuda.synchronize()
            reps = 2000
            t0 = time.perf_counter()
            for _ in range(reps):
                a * b
            host_us = (time.perf_counter() - t0) / reps * 1e6      # launches queue up: this is issue cost when < kernel time
            torch.cuda.synchronize()
            # the same call on a 4096-item batch (kernel ~2 us): the GPU never limits, so this is the HOST cost of one
            # eager operator application (python dispatch + result allocation + plan launch in csrc/fastcall.cpp)
            sa, sb = device_mv(pga, S.even_grade(), 4096, 21, rank), device_mv(pga, S.even_grade(), 4096, 22, rank)
            for _ in range(10):
                sa * sb
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                sa * sb
            host_small_us = (time.perf_counter() - t0) / reps * 1e6
            torch.cuda.synchronize()

            def twenty(a, b):
                for _ in range(20):
                    o = a * b
                return o
            g = pga.capture(twenty, a, b)
            ms_graph = time_steps(g.replay, 10, 3) / 20
            extra = {'eager_call_us_wall': host_us, 'host_us_per_eager_call': host_small_us, 'ms_in_cuda_graph': ms_graph,
                     'frac_in_cuda_graph': 96 * n / ms_graph / 1e6 / peak,
                     'l2_note': '96 MB working set fits the 126 MB L2: fractions above 1.0 of the HBM copy peak are L2 hits'}
        if faithful:
            fctx = B200Context((3, 0, 1), dtype=torch.float32, mode='faithful')
            fa, fb = fctx.multivector(values=a.values, subspace=a.subspace), fctx.multivector(values=b.values, subspace=b.subspace)
            extra['faithful_ms'] = time_steps(lambda: fa * fb, 20, 3)
        out.append(entry('cfg1', f'config 1: 3D PGA motor*motor, fp32, batch 2^{int(np.log2(n))}, eager call', n, ms, 96, peak,
                         flops_per_item=88, parity_ok=bool(err <= TOL['cfg1']), parity_err=err, parity_tol=TOL['cfg1'], **extra))
        del a, b
    return out


def run_cfg3(torch, B200Context, rank, peak, faithful=True):
    """configs[2]: bivector -> exp -> normalized -> motor_log round trip, 2^28, fp32 and fp64, ONE kernel."""
    from oracle import numga_oracle as O
    out = []
    issue = {}
    ip = os.path.join(ROOT, 'profiles', 'r02_issue_slots.json')
    if os.path.exists(ip):
        issue = json.load(open(ip))
    for dtype, tag, ub in ((torch.float32, 'f32', 48), (torch.float64, 'f64', 96)):
        npd = np.float32 if tag == 'f32' else np.float64
        c = B200Context((3, 0, 1), dtype=dtype)
        f = c.jit(lambda x: x.exp().normalized().motor_log())
        hb = (0.5 * np.random.default_rng(103).standard_normal((PARITY_N, 6)))
        with np.errstate(all='ignore'):
            want64 = O.context((3, 0, 1), np.float64).multivector('bivector', hb).exp().normalized().motor_log().values
            want = O.context((3, 0, 1), npd).multivector('bivector', hb.astype(npd)).exp().normalized().motor_log().values
        got = f(c.multivector.bivector(hb.astype(npd))).numpy()
        # gated against the reference at the SAME precision; the errors against the fp64 result are reported next to it
        e_gpu, e_ref, e_vs_ref = rel_err(got, want64), rel_err(want, want64), rel_err(got, want)
        tol = TOL['cfg3_' + tag]
        n = 1 << 28
        bv = device_mv(c, c.subspace.bivector(), n, 13, rank, scale=0.5)
        ms = time_steps(lambda: f(bv), 5, 3)
        prog = c.engine.cache[next(k for k in c.engine.cache if k[0][0] == 'jit')].program
        extra = {}
        if faithful and tag == 'f32':
            fc = B200Context((3, 0, 1), dtype=dtype, mode='faithful')
            ff = fc.jit(lambda x: x.exp().normalized().motor_log())
            fb = fc.multivector(values=bv.values, subspace=bv.subspace)
            extra['faithful_ms'] = time_steps(lambda: ff(fb), 3, 2)
            extra['faithful_bit_exact_vs_oracle'] = bool(np.array_equal(ff(fc.multivector.bivector(hb.astype(npd))).numpy(), want))
        e = entry('cfg3_' + tag, f'config 3: exp->normalized->motor_log round trip, {tag}, batch 2^28, one fused kernel', n, ms, ub, peak,
                  flops_per_item=prog.info['n_instr'], tflops=prog.info['n_instr'] * n / ms / 1e9,
                  parity_ok=bool(e_vs_ref <= tol), parity_err=e_vs_ref, parity_err_vs_fp64=e_gpu, reference_err_vs_fp64=e_ref,
                  parity_tol=tol, **extra)
        e['roofline']['bound'] = 'fp32 issue' if tag == 'f32' else 'fp64 pipe'
        e['roofline']['compute_bound'] = True
        e['roofline']['issue_slot_utilisation_ncu'] = issue.get('cfg3_' + tag)
        out.append(e)
        # supplementary: the same round trip with the opt-in closed forms (numga_b200.optimized: the reference's optional
        # extension/optimized.py for exp / normalized, plus the inverse of its exp as the logarithm) -- one HBM-bound kernel
        import numga_b200.optimized as opt
        opt.enable()
        try:
            g = c.jit(lambda x: x.exp().normalized().motor_log())
            got_c = g(c.multivector.bivector(hb.astype(npd))).numpy()
            ms_c = time_steps(lambda: g(bv), 5, 3)
            n_c = max(v.program.info['n_instr'] for k, v in c.engine.cache.items() if k[0][0] == 'jit')
        finally:
            opt.disable()
        tol_c = 1e-5 if tag == 'f32' else 1e-12
        ec = entry('cfg3_' + tag + '_closed_form', f'config 3 (supplementary): the same round trip with the OPT-IN closed forms of '
                   f'numga_b200.optimized (3D PGA only), {tag}, batch 2^28, one kernel', n, ms_c, ub, peak, flops_per_item=n_c,
                   parity_ok=bool(rel_err(got_c, want64) <= tol_c), parity_err_vs_fp64=rel_err(got_c, want64),
                   reference_err_vs_fp64=e_ref, parity_tol=tol_c)
        ec['note'] = ('not the default dispatch: closed forms replace the reference\'s bisection algorithms (its own optional '
                      'extension/optimized.py does so for exp / normalized; the logarithm is the inverse of that exp); gated against '
                      'the fp64 oracle, where it is MORE accurate than the reference\'s same-precision bisection (reference_err_vs_fp64)')
        out.append(ec)
        del bv
        torch.cuda.empty_cache()
    return out


def run_cfg4(torch, B200Context, rank, peak, faithful=True):
    """configs[3]: CGA (4,1,0) full x full geometric product, 2^26, fp32, SIMT (tensor-core note in DESIGN.md)."""
    from oracle import numga_oracle as O
    cga = B200Context((4, 1, 0), dtype=torch.float32)
    F = cga.subspace.full()
    rng = np.random.default_rng(104)
    hx, hy = rng.standard_normal((PARITY_N, 32)).astype(np.float32), rng.standard_normal((PARITY_N, 32)).astype(np.float32)
    oc = O.context((4, 1, 0), np.float32)
    want = (oc.multivector('multivector', hx) * oc.multivector('multivector', hy)).values
    err = rel_err((cga.multivector(values=hx, subspace=F) * cga.multivector(values=hy, subspace=F)).numpy(), want)
    n = 1 << 26
    x, y = device_mv(cga, F, n, 14, rank), device_mv(cga, F, n, 15, rank)
    ms = time_steps(lambda: x * y, 10, 3)
    extra = {}
    if faithful:
        fc = B200Context((4, 1, 0), dtype=torch.float32, mode='faithful')
        fx, fy = fc.multivector(values=x.values, subspace=F), fc.multivector(values=y.values, subspace=F)
        extra['faithful_ms'] = time_steps(lambda: fx * fy, 5, 2)
    e = entry('cfg4', 'config 4: CGA (4,1,0) full*full geometric product, fp32, batch 2^26 (SIMT, 1024 FMA/item)', n, ms, 384, peak,
              flops_per_item=2016, tflops=2016 * n / ms / 1e9, parity_ok=bool(err <= TOL['cfg4']), parity_err=err,
              parity_tol=TOL['cfg4'], **extra)
    e['tensor_core_comparison'] = ('not run: every K[i,:,:] of the [32,32,32] table is a signed permutation, so a GEMM form computes '
                                   '32x the flops (65.5 kflop/item => 1.1 PFLOP/s TF32 at the HBM rate, above the TF32 peak) and TF32 '
                                   'misses the 1e-5 gate; see DESIGN.md "Tensor cores"')
    del x, y
    torch.cuda.empty_cache()
    return [e]


def physics_state(ctx, n_chains, rank, seed=16, host_rate=None):
    from numga_b200.examples.physics import replicate_chains, setup_bodies
    bodies, sets = setup_bodies(ctx, n_bodies=12)
    big, bsets = replicate_chains(bodies, sets, n_chains)
    if host_rate is not None:
        big = big.copy(rate=big.rate + ctx.multivector.bivector(host_rate))
    else:
        big = big.copy(rate=big.rate + device_mv(ctx, ctx.subspace.bivector(), 12 * n_chains, seed, rank, scale=0.3))
    return big, bsets


def run_cfg5(torch, B200Context, rank, peak, n_target=1 << 24, unroll=10, faithful=True):
    """configs[4]: rigid-body XPBD core step (numga/examples/physics/core.py Body.integrate, two colour sets) on ~2^24
    bodies = independent 12-body chains.  Two kernels of the same physics are timed:
      * `ChainStep`: ONE chain-resident kernel (state once through HBM per launch; `unroll` sub-steps per launch as the
        reference's run_chain_numpy.py:26-30 drives it);
      * `FusedStep`: six kernels per sub-step (round-1 arrangement), for comparison."""
    from oracle import physics_oracle as P
    from numga_b200 import runtime as rt
    from numga_b200.examples import physics as X
    out = []
    issue = {}
    ip = os.path.join(ROOT, 'profiles', 'r02_issue_slots.json')
    if os.path.exists(ip):
        issue = json.load(open(ip))
    for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        npd = np.float32 if tag == 'f32' else np.float64
        es = 4 if tag == 'f32' else 8
        c = B200Context('x+y+z+w0', dtype=dtype)
        # ---- parity of the timed kernels on 2^16 bodies against the oracle port of the reference's sub-step -------
        pc = PARITY_N // 12
        hrate = (0.3 * np.random.default_rng(105).standard_normal((12 * pc, 6))).astype(npd)
        small, ssets = physics_state(c, pc, rank, host_rate=hrate)
        arrays = physics_arrays(tag)
        ob, osets = P.from_arrays(arrays, npd, n_chains=pc, rate=np.tile(arrays['init/rate'], (pc, 1)).astype(npd) + hrate)
        with np.errstate(all='ignore'):
            want1 = ob.integrate(PHYSICS_DT, osets)
            want2 = want1.integrate(PHYSICS_DT, osets)
        mt, rt_tol = TOL['cfg5_' + tag]
        steppers = {'FusedStep': X.FusedStep(small, ssets)}
        if hasattr(X, 'ChainStep'):
            steppers['ChainStep'] = X.ChainStep(small, ssets, unroll=2)
            if tag == 'f32':
                steppers['ChainStep(exploit_zeros=True)'] = X.ChainStep(small, ssets, unroll=2, exploit_zeros=True)
        parity = {}
        for name, st in steppers.items():
            if name.startswith('ChainStep'):
                got2 = st.integrate(small, PHYSICS_DT)
            else:
                got2 = st.integrate(st.integrate(small, PHYSICS_DT), PHYSICS_DT)
            em, er = rel_err(got2.motor.numpy(), want2.motor.values), rel_err(got2.rate.numpy(), want2.rate.values)
            parity[name] = {'motor_err': em, 'rate_err': er, 'ok': bool(em <= 2 * mt and er <= 2 * rt_tol), 'sub_steps': 2,
                            'tol': [2 * mt, 2 * rt_tol]}
        del small, ssets, steppers
        # ---- timing ---------------------------------------------------------------------------------------------------
        n_chains = n_target // 12
        n_bodies = 12 * n_chains
        big, bsets = physics_state(c, n_chains, rank)
        n_constraints = sum(int(s.body_idx.shape[1]) for s in bsets)
        state_once = (92 + 14) * es + (2 * 4 + 2 * 36 + 1) * es * n_constraints / n_bodies + 2 * 8 * n_constraints / n_bodies
        fs = X.FusedStep(big, bsets)
        state = [big]

        def step():
            state[0] = fs.integrate(state[0], PHYSICS_DT)
        l0 = rt.launch_count()
        step()
        launches = rt.launch_count() - l0
        ms6 = time_steps(step, 10, 3)
        words = {}
        for k, v in c.engine.cache.items():
            if k[0][0] == 'jit':
                live_in = sum(1 for op, a, b, _ in v.ir['instr'] if op == rt.OP_INPUT and v.program.in_mode[a] != rt.BROADCAST)
                words[k[0][1].__name__] = live_in + sum(v.ir['out_width'])
        split_bytes = ((words['pre'] + words['post']) * es * n_bodies +
                       ((words['apply'] + words['v_apply']) * es + 2 * 2 * 8) * n_constraints) / n_bodies
        e6 = entry('cfg5_' + tag, f'config 5: rigid-body XPBD sub-step, {tag}, {n_bodies} bodies in {n_chains} chains, FusedStep: '
                   f'{launches} kernels per sub-step', n_bodies, ms6, state_once, peak, unit='body-steps/s',
                   parity_ok=parity['FusedStep']['ok'], parity=parity['FusedStep'], kernel_launches_per_step=launches)
        e6['roofline']['bytes_per_item_note'] = 'SURVEY.md 8d state-once algorithmic bytes (each state word once per sub-step)'
        e6['roofline_multi_kernel_convention'] = {'bytes_per_item': split_bytes, 'achieved': split_bytes * n_bodies / ms6 / 1e6,
                                                  'frac': split_bytes * n_bodies / ms6 / 1e6 / peak,
                                                  'note': 'live inputs + outputs of the six kernels, incl. gathered inertia matrices'}
        out.append(e6)
        if hasattr(X, 'ChainStep'):
            variants = [(1, False), (unroll, False)] + ([(1, True), (unroll, True)] if tag == 'f32' else [])
            for u, zeros in variants:
                cs = X.ChainStep(big, bsets, unroll=u, exploit_zeros=zeros)
                state = [big]

                def cstep():
                    state[0] = cs.integrate(state[0], PHYSICS_DT)
                l0 = rt.launch_count()
                cstep()
                launches = rt.launch_count() - l0
                ms = time_steps(cstep, 5 if u > 1 else 10, 3)
                label = 'ChainStep' + ('(exploit_zeros=True)' if zeros else '')
                e = entry('cfg5_' + tag + ('_zeros' if zeros else ''),
                          f'config 5: rigid-body XPBD, {tag}, {n_bodies} bodies, {label}: ONE chain-resident kernel, '
                          f'{u} sub-step(s) per launch', n_bodies * u, ms, state_once / u, peak, unit='body-steps/s',
                          parity_ok=parity[label]['ok'], parity=parity[label], kernel_launches_per_step=launches,
                          sub_steps_per_launch=u, bodies=n_bodies)
                e['roofline']['dram_bytes_per_body_ncu'] = issue.get(f"cfg5_{tag}_dram_bytes_per_body") if not zeros else None
                e['anchor_maps'] = ('re-evaluated in registers from the anchors (ChainStep(recompute_maps=True), the default) instead of '
                                    'streaming the 2 x 24-word tables per constraint' if cs.recompute_maps else 'streamed from the tabulated anchors_map')
                moved = e['roofline']['dram_bytes_per_body_ncu']
                if moved and u == 1:
                    # what this kernel really moves (ncu, profiles/r02_chain_*_final_ncu_full.csv): fewer bytes than SURVEY's
                    # state-once figure, because the anchor maps are not read and 12 inertia entries are structurally dead
                    e['roofline']['hbm_frac_of_bytes_moved'] = moved * n_bodies / ms / 1e6 / peak
                    e['roofline']['note'] = (f'frac is quoted on SURVEY 8d\'s {state_once:.1f} B/body as the contract asks; the kernel moves only '
                                             'dram_bytes_per_body_ncu, so HBM is ~half busy and the kernel is bound by issue slots / latency '
                                             '(issue_slot_utilisation_ncu), not by HBM')
                if not zeros:
                    e['roofline']['issue_slot_utilisation_ncu'] = issue.get(f"cfg5_{tag}_{'one_substep' if u == 1 else 'ten_substeps'}_per_launch")
                    if u > 1:
                        e['roofline']['bound'], e['roofline']['compute_bound'] = 'issue slots', True
                if zeros:
                    e['note'] = ('supplementary: entries of the inertia / inverse-inertia / anchor-map tables that are exactly zero for '
                                 'EVERY body (found by inspecting the tables at set-up; 6 of 36 in the inverse inertia of this example) '
                                 'are compiled out, so fewer bytes than the 735.7 B/body of the dense treatment cross HBM; the roofline '
                                 'fraction is still quoted on the dense figure and can exceed what a dense kernel could reach')
                e['roofline']['bytes_per_item_note'] = ('state-once bytes per launch / sub-steps per launch: the state crosses HBM once '
                                                        'per launch' + ('; compute-bound at this unroll' if u > 1 else ''))
                e['finite'] = bool(torch.isfinite(state[0].motor.values).all().item() and torch.isfinite(state[0].rate.values).all().item())
                out.append(e)
                del cs
        del big, bsets, fs, state
        torch.cuda.empty_cache()
    return out


def extra_configs(torch, B200Context, rank, peak):
    """--extra: supplementary measurements (not BASELINE configs)."""
    out = []
    import numga_b200.optimized as opt
    c = B200Context((3, 0, 1), dtype=torch.float32)
    n = 1 << 28
    bv = device_mv(c, c.subspace.bivector(), n, 13, rank, scale=0.5)
    for tag, on in (('generic bisection (default dispatch)', False), ('opt-in closed form (numga_b200.optimized)', True)):
        (opt.enable if on else opt.disable)()
        f = c.jit(lambda x: x.exp())
        ms = time_steps(lambda: f(bv), 5, 2)
        out.append(entry('extra', f'bivector exp only, fp32, batch 2^28, {tag}', n, ms, 56, peak))
    opt.disable()
    del bv
    out.append(training_step_config(torch, B200Context, rank, peak))
    out.extend(generic_physics_configs(torch, B200Context, rank, peak, modes=('eager',)))
    return out


def generic_physics_configs(torch, B200Context, rank, peak, n_target=1 << 22, modes=('eager', 'lazy', 'lazy+graph')):
    """The rigid-body step written exactly like the reference (Body.integrate, one multivector method after the other):
    eagerly (one kernel per method), with automatic fusion (B200Context(lazy=True): 7 kernels per sub-step, host-bound by
    the python recording), and with that lazy step captured ONCE into a CUDA graph (B200Context.capture) and replayed."""
    from numga_b200 import runtime as rt
    out = []
    n_chains = n_target // 12
    for mode in modes:
        ctx = B200Context('x+y+z+w0', dtype=torch.float32, lazy=mode.startswith('lazy'))
        big, bsets = physics_state(ctx, n_chains, rank)
        big.rate.values
        state = [big]

        def advance(b):
            new = b.integrate(PHYSICS_DT, bsets)
            new.motor.values, new.rate.values
            return new
        if mode == 'lazy+graph':
            static = big.copy(motor=big.motor.copy(), rate=big.rate.copy())          # the graph's input buffers
            graph = ctx.capture(advance, static)
            launches = graph.n_kernels

            def step():
                new = graph.replay()
                static.motor.values.copy_(new.motor.values)                           # feed the result back (2 small copies)
                static.rate.values.copy_(new.rate.values)
        else:
            def step():
                state[0] = advance(state[0])
            l0 = rt.launch_count()
            step()
            launches = rt.launch_count() - l0
        ms = time_steps(step, 5, 2)
        n = 12 * n_chains
        label = {'eager': 'eager: one generated kernel per multivector method', 'lazy': 'lazy=True: automatic fusion, python recording every step',
                 'lazy+graph': 'lazy=True captured once into a CUDA graph (ctx.capture) and replayed'}[mode]
        out.append({'id': 'cfg5_reference_style', 'config': f'reference-style Body.integrate (unchanged formulas; {label}; {launches} generated kernels per '
                    f'sub-step), fp32, {n} bodies', 'items': n, 'ms': ms, 'items_per_s': n / ms * 1e3, 'unit': 'body-steps/s',
                    'kernel_launches_per_step': launches,
                    'roofline': {'bound': 'hbm', 'bytes_per_item': 735.67, 'achieved': 735.67 * n / ms / 1e6, 'peak': peak, 'unit': 'GB/s',
                                 'frac': 735.67 * n / ms / 1e6 / peak}})
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
    e = entry('extra', f'autograd training step (fit one motor to 2^{int(np.log2(n))} points: forward + torch sum + generated VJP '
              'kernel with in-kernel reduction)', n, ms, nbytes / n, peak)
    e['grad_finite'] = bool(torch.isfinite(b.grad).all().item())
    return e


def host_roundtrip_control(rt, local, n_in_bytes, n_out_bytes, hin, hout, steps, dist):
    """Raw-copy control of the host path: the same pinned buffers, the same chunking and streams as
    ngb200_program_run_host, H2D of the inputs and D2H of the outputs, NO kernel; all ranks concurrently."""
    import ctypes
    if not hasattr(rt.lib, 'ngb200_host_copy_control'):
        return None
    fn = rt.lib.ngb200_host_copy_control
    rt.check(fn(hin, n_in_bytes, hout, n_out_bytes, 0, 0, local))
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        rt.check(fn(hin, n_in_bytes, hout, n_out_bytes, 0, 0, local))
    return (time.perf_counter() - t0) / steps


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

    # ---- parity of the timed kernel on EVERY rank: first and last 2^16 points of this rank's slice against the oracle
    # (the Philox stream is deterministic per offset = rank * n; the inputs are read back, the oracle runs on the host)
    from oracle import numga_oracle as O
    oc = O.context((3, 0, 1), np.float32)
    out_dev = sandwich(R, p)
    perr = 0.0
    for sl in (slice(0, PARITY_N), slice(n - PARITY_N, n)):
        hp = p.values[sl].cpu().numpy()
        want = (oc.multivector('even_grade', R.numpy()) >> oc.multivector('antivector', hp)).values
        perr = max(perr, rel_err(out_dev.values[sl].cpu().numpy(), want))
    parity_ok = perr <= TOL['cfg2']
    if dist is not None:
        t = torch.tensor([0.0 if parity_ok else 1.0, perr], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parity_ok_all, perr_all = bool(t[0].item() == 0.0), float(t[1].item())
    else:
        parity_ok_all, perr_all = bool(parity_ok), perr
    fctx = B200Context((3, 0, 1), dtype=torch.float32, device=f'cuda:{local}', mode='faithful')
    fsand = fctx.jit(lambda r, x: r >> x)
    fR, fp = fctx.multivector(values=R.values, subspace=R.subspace), fctx.multivector(values=p.values, subspace=p.subspace)
    faithful_ms = time_steps(lambda: fsand(fR, fp), 5, 2)
    del out_dev

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
    check = sandwich(R, p).values[:1 << 16].T.cpu().numpy()
    e2e_ok = bool(np.array_equal(check, host_q[:, :1 << 16]))
    ctrl_s = host_roundtrip_control(rt, local, 16 * e2e_n, 16 * e2e_n, hin, hout, e2e_steps, dist)   # (overwrites host_q)
    if dist is not None:
        t = torch.tensor([e2e_s, ctrl_s or 0.0], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, ctrl_s = float(t[0].item()), (float(t[1].item()) if ctrl_s else None)
    e2e = {'value': world * e2e_n / e2e_s, 'unit': 'points/s', 'h2d_bytes_per_step': 16 * e2e_n + 32,
           'd2h_bytes_per_step': 16 * e2e_n, 'steps': e2e_steps, 'points_per_gpu': e2e_n,
           'matches_device_result': e2e_ok, 'numa_local_cpus': numa,
           'path': 'ngb200_program_run_host (pinned host SoA buffers, 128 MB chunks, H2D/kernel/D2H pipelined over 8 streams)'}
    if ctrl_s:
        e2e['host_ceiling_points_s'] = world * e2e_n / ctrl_s
        e2e['frac_of_host_ceiling'] = ctrl_s / e2e_s
        e2e['host_ceiling_how'] = ('ngb200_host_copy_control: same pinned buffers, chunks and streams, H2D of the inputs + D2H of the '
                                   'outputs, no kernel, all ranks concurrently; pinned GB/s per GPU = '
                                   f'{32 * e2e_n / ctrl_s / 1e9:.1f}')
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
        del src, dst
    del p
    torch.cuda.empty_cache()

    # ---- multi-GPU: the physics configuration sharded by chain (no exchange), every rank its own chains ------------
    physics_multi = None
    if dist is not None and not args.no_configs:
        from numga_b200.examples import physics as X
        c5 = B200Context('x+y+z+w0', dtype=torch.float32, device=f'cuda:{local}')
        big, bsets = physics_state(c5, (1 << 24) // 12, rank)
        stepper = X.ChainStep(big, bsets, unroll=1) if hasattr(X, 'ChainStep') else X.FusedStep(big, bsets)
        st = [big]

        def pstep():
            st[0] = stepper.integrate(st[0], PHYSICS_DT)
        ms5 = time_steps(pstep, 10, 3, dist)
        nb = 12 * ((1 << 24) // 12)
        physics_multi = {'config': 'config 5 sharded by chain, fp32, 2^24 bodies per GPU, ' + type(stepper).__name__,
                         'body_steps_per_s': world * nb / ms5 * 1e3, 'ms': ms5, 'bodies_per_gpu': nb}
        del big, bsets, st, stepper
        torch.cuda.empty_cache()

    line = None
    if rank == 0:
        import hashlib
        source_sha1 = hashlib.sha1(prog.source.encode()).hexdigest()[:16]       # staleness check of the cited ncu capture
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, 'profiles', 'r02_sandwich_traffic.json')
        if not os.path.exists(tp):
            tp = os.path.join(ROOT, 'profiles', 'r01_sandwich_traffic.json')
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic, traffic_src = tj.get('dram_bytes_per_launch'), {'file': os.path.relpath(tp, ROOT), 'commit': tj.get('commit'),
                                                                      'kernel_source_sha1': tj.get('kernel_source_sha1')}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'points/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': 'BASELINE.json configs[1]: 3D PGA (3,0,1) motor sandwich R >> p, one broadcast motor, fp32',
                       'points_per_gpu': n, 'layout': 'structure-of-arrays [4, n] in HBM', 'parallelism': f'batch-sharded x{world}, no collective',
                       'l2': f'inputs+outputs {2 * 16 * n / 1e9:.1f} GB per GPU per step >> 126 MB L2 (no flush needed)'},
            'clocks': clock_summary, 'gpu_launches': launches, 'e2e': e2e,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': BYTES_PER_POINT * n, 'kernel_ms': ms,
                         'kernel_source_sha1': source_sha1, 'traffic_is_current': bool(traffic_src and traffic_src.get('kernel_source_sha1') == source_sha1)},
            'parity_ok_all_ranks': parity_ok_all, 'parity_err_max_over_ranks': perr_all, 'parity_tol': TOL['cfg2'],
            'parity_how': 'every rank: first and last 2^16 points of its slice, timed kernel vs oracle port on the host',
            'faithful_ms_per_step': faithful_ms,
        }
        if gather:
            line['gather'] = gather
        if physics_multi:
            line['physics_sharded'] = physics_multi
    if world == 1 and not args.no_configs:
        cfgs = []
        cfgs += run_cfg1(torch, B200Context, rank, peak)
        cfgs.append(entry('cfg2', 'config 2 (headline, see top level): R >> p, 2^28 points, broadcast motor, fp32', n, ms, 32, peak,
                          unit='points/s', flops_per_item=28, parity_ok=parity_ok_all, parity_err=perr_all, parity_tol=TOL['cfg2'],
                          faithful_ms=faithful_ms))
        cfgs += run_cfg3(torch, B200Context, rank, peak)
        cfgs += run_cfg4(torch, B200Context, rank, peak)
        cfgs += run_cfg5(torch, B200Context, rank, peak)
        cfgs += generic_physics_configs(torch, B200Context, rank, peak, modes=('lazy', 'lazy+graph'))
        if args.extra:
            cfgs += extra_configs(torch, B200Context, rank, peak)
        line['configs'] = cfgs
    if rank == 0 and world == 1 and not args.no_cpu:
        farm = CpuFarm()
        units = {'cfg1': 'products/s', 'cfg2': 'points/s', 'cfg3_f32': 'products/s', 'cfg3_f64': 'products/s', 'cfg4': 'products/s',
                 'cfg5_f32': 'body-steps/s', 'cfg5_f64': 'body-steps/s'}
        wanted = list(units) if not args.no_configs else ['cfg2']
        base = {cfg: farm.baselines(cfg, units[cfg]) for cfg in wanted}
        farm.close()
        b2 = base['cfg2']
        line['cpu_baseline'] = {'value': b2['einsum_allcores'], 'unit': 'points/s', 'cores': b2['cores'], 'kind': 'port',
                                'sample': f"{b2['cores']} pinned processes x 2^22 points per step (NumpyEinsumOperator port), median step",
                                'einsum_1core': b2['einsum_1core'], 'sparse_1core': b2['sparse_1core'], 'sparse_allcores': b2['sparse_allcores']}
        for e in line.get('configs', []):
            if e['id'] in base:
                e['cpu_baseline'] = base[e['id']]
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
    ap.add_argument('--no-configs', action='store_true', help='headline configuration only (skip BASELINE configs 1, 3, 4, 5)')
    ap.add_argument('--configs', action='store_true', help='(default now) kept for compatibility')
    ap.add_argument('--extra', action='store_true', help='also run the supplementary measurements')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline legs')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
