"""Sweep the host-buffer path (ngb200_program_run_host) over chunk size / stream count at the headline size."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numga_b200 import B200Context, runtime as rt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
ctx = B200Context((3, 0, 1))
S = ctx.subspace
R = ctx.multivector.motor(np.random.default_rng(1).standard_normal(8).astype(np.float32)).normalized()
f = ctx.jit(lambda r, x: r >> x)
f(R, ctx.multivector.antivector(torch.zeros(64, 4, device='cuda')))
prog = ctx.engine.cache[next(k for k in ctx.engine.cache if k[0][0] == 'jit')].program
hin, hout, hR = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
rt.check(rt.lib.ngb200_host_alloc(ctypes.byref(hin), 16 * n))
rt.check(rt.lib.ngb200_host_alloc(ctypes.byref(hout), 16 * n))
rt.check(rt.lib.ngb200_host_alloc(ctypes.byref(hR), 32))
host_p = np.ctypeslib.as_array(ctypes.cast(hin, ctypes.POINTER(ctypes.c_float)), shape=(4, n))
host_R = np.ctypeslib.as_array(ctypes.cast(hR, ctypes.POINTER(ctypes.c_float)), shape=(8,))
host_R[:] = R.numpy()
host_p[:] = 1.0
ins = [(hR.value, 1, 0), (hin.value, n, 1)]
outs = [(hout.value, n, 1)]
for mb, streams in [(64, 4), (16, 4), (32, 4), (128, 4), (256, 4), (32, 2), (32, 3), (32, 8), (64, 8), (8, 8), (512, 2)]:
    os.environ['NGB200_HOST_CHUNK_MB'] = str(mb)
    os.environ['NGB200_HOST_STREAMS'] = str(streams)
    prog.run_host(ins, outs, n)
    t0 = time.perf_counter()
    for _ in range(3):
        prog.run_host(ins, outs, n)
    dt = (time.perf_counter() - t0) / 3
    print(f'chunk {mb:4d} MB streams {streams}: {dt*1e3:8.1f} ms  {n/dt/1e9:6.3f} G points/s  {32*n/dt/1e9:6.1f} GB/s PCIe (both directions)', flush=True)
# plain copies for reference: H2D alone, D2H alone, both concurrently
d_in = torch.empty(4 * n, dtype=torch.float32, device='cuda'); d_out = torch.empty(4 * n, dtype=torch.float32, device='cuda')
hin_t = torch.from_numpy(host_p.reshape(-1)); hout_np = np.ctypeslib.as_array(ctypes.cast(hout, ctypes.POINTER(ctypes.c_float)), shape=(4 * n,)); hout_t = torch.from_numpy(hout_np)
for name, fn in (('H2D only', lambda: d_in.copy_(hin_t, non_blocking=True)), ('D2H only', lambda: hout_t.copy_(d_out, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'{name}: {16*n/dt/1e9:.1f} GB/s  (pinned={hin_t.is_pinned()})')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d_in.copy_(hin_t, non_blocking=True)
with torch.cuda.stream(s2): hout_t.copy_(d_out, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f'H2D + D2H concurrently: {32*n/dt/1e9:.1f} GB/s total -> best possible e2e {n/dt/1e9:.3f} G points/s')
