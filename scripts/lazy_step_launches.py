"""The reference-style physics step under lazy=True, a few times: target for an ncu launch list
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numga_b200 import B200Context
from bench import physics_state, PHYSICS_DT
ctx = B200Context('x+y+z+w0', dtype=torch.float32, lazy=True)
big, bsets = physics_state(ctx, (1 << 22) // 12, 0)
big.rate.values
state = [big]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    new = state[0].integrate(PHYSICS_DT, bsets)
    new.motor.values, new.rate.values
    state[0] = new
torch.cuda.synchronize()
print('done')
