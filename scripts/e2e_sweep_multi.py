"""Host-buffer path under contention: every rank runs ngb200_program_run_host on its own 2^28-point slice at the same
time (torchrun, one rank per GPU), for several (chunk size, stream count) settings, next to the raw-copy control
(ngb200_host_copy_control).  Prints whole-job G points/s (max time over ranks).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/e2e_sweep_multi.py"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from numga_b200 import B200Context, runtime as rt

rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 1 << 28
ctx = B200Context((3, 0, 1), device=f'cuda:{local}')
R = ctx.multivector.motor(np.random.default_rng(1).standard_normal(8).astype(np.float32)).normalized()
f = ctx.jit(lambda r, x: r >> x)
f(R, ctx.multivector.antivector(torch.zeros(64, 4, device=f'cuda:{local}')))
prog = ctx.engine.cache[next(k for k in ctx.engine.cache if k[0][0] == 'jit')].program
hin, hout, hR = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
for h, b in ((hin, 16 * n), (hout, 16 * n), (hR, 32)):
    rt.check(rt.lib.ngb200_host_alloc(ctypes.byref(h), b))
np.ctypeslib.as_array(ctypes.cast(hR, ctypes.POINTER(ctypes.c_float)), shape=(8,))[:] = R.numpy()
np.ctypeslib.as_array(ctypes.cast(hin, ctypes.POINTER(ctypes.c_float)), shape=(4, n))[:] = 1.0
ins, outs = [(hR.value, 1, 0), (hin.value, n, 1)], [(hout.value, n, 1)]


def timed(fn, reps=3):
    fn()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


for mb, streams in [(128, 4), (64, 4), (32, 4), (128, 8), (64, 8), (32, 8), (16, 8), (64, 16), (32, 16), (256, 4)]:
    os.environ['NGB200_HOST_CHUNK_MB'], os.environ['NGB200_HOST_STREAMS'] = str(mb), str(streams)
    dt = timed(lambda: prog.run_host(ins, outs, n, device=local))
    dc = timed(lambda: rt.check(rt.lib.ngb200_host_copy_control(hin, 16 * n, hout, 16 * n, 0, 0, local)))
    if rank == 0:
        print(f'e2e x{world} chunk {mb:4d} MB streams {streams:2d}: pipeline {world * n / dt / 1e9:6.3f} G points/s, raw copies '
              f'{world * n / dc / 1e9:6.3f} G points/s, ratio {dc / dt:.3f}', flush=True)
if world > 1:
    dist.destroy_process_group()
