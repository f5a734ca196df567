"""Host-side profile (cProfile) of the reference-style physics step under lazy=True on the GPU. Usage: python scripts/lazy_host_profile.py"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numga_b200 import B200Context
from bench import physics_state, PHYSICS_DT
ctx = B200Context('x+y+z+w0', dtype=torch.float32, lazy=True)
big, bsets = physics_state(ctx, (1 << 22) // 12, 0)
big.rate.values
state = [big]
def step():
    new = state[0].integrate(PHYSICS_DT, bsets)
    new.motor.values, new.rate.values
    state[0] = new
for _ in range(5): step()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize(); print('ms per step', (time.perf_counter() - t) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumtime').print_stats(45)
