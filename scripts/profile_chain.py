"""Run the chain-resident physics kernel (ChainStep) a few times: target for ncu / quick timing sweeps.
Usage: profile_chain.py [f32|f64] [log2 bodies] [unroll] [bodies_per_tile]   (env knobs NGB200_* apply)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numga_b200 import B200Context
from numga_b200.examples.physics import ChainStep
from bench import physics_state, time_steps, PHYSICS_DT

tag = sys.argv[1] if len(sys.argv) > 1 else 'f32'
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 22
unroll = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bpt = int(sys.argv[4]) if len(sys.argv) > 4 else None
tpt = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ctx = B200Context('x+y+z+w0', dtype=torch.float32 if tag == 'f32' else torch.float64)
big, bsets = physics_state(ctx, (1 << logn) // 12, 0)
pairs = os.environ.get('CHAIN_PAIRS', '0') == '1'          # two threads per constraint (ChainStep(pairs=True))
cs = ChainStep(big, bsets, unroll=unroll, bodies_per_tile=bpt, threads_per_tile=tpt, pairs=pairs)
state = [big]


def step():
    state[0] = cs.integrate(state[0], PHYSICS_DT)
ms = time_steps(step, 5, 2)
n = big.motor.shape[0]
knobs = {k: v for k, v in os.environ.items() if k.startswith('NGB200') and 'CACHE' not in k}
print(f'chain {tag} 2^{logn} unroll {unroll} bpt {bpt} tpt {tpt} pairs {int(pairs)} {knobs}: {ms:.3f} ms, {n * unroll / ms / 1e6:.3f} G body-steps/s, info {cs.compiled.program.info}', flush=True)
