"""Time one training-style step of the headline workload: fit ONE motor to N point correspondences
(forward R >> p, per-item squared error, backward through the generated VJP kernel with an in-kernel batch reduction)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numga_b200 import B200Context, runtime as rt
from bench import device_mv, time_steps

n = 1 << int(os.environ.get('LOGN', '26'))
ctx = B200Context((3, 0, 1)); S = ctx.subspace
p = device_mv(ctx, S.antivector(), n, 2, 0)
target = ctx.multivector.bivector(np.array([0.3, -0.2, 0.1, 0.5, 0.2, -0.4], np.float32)).exp()
q = target >> p
b = torch.zeros(6, device='cuda', requires_grad=True)
def sq_err(r, x, y):
    """per-item sum of squared component differences, as a scalar multivector (computed inside the same kernel)"""
    d = ((r >> x) - y).values
    return ctx.multivector.scalar(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2] + d[..., 3] * d[..., 3])
err = ctx.jit(sq_err)

def step():
    b.grad = None
    R = ctx.multivector(values=b, subspace=S.bivector()).exp()
    loss = err(R, p, q).values.sum()
    loss.backward()
    return loss
l0 = rt.launch_count(); loss = step(); torch.cuda.synchronize(); launches = rt.launch_count() - l0
ms = time_steps(step, 10, 3)
fwd_bytes, bwd_bytes = (16 + 16 + 4) * n, (16 + 16 + 4) * n      # bwd: x, y, cotangent in; 8 reduced floats out
print(f'train step N=2^{int(np.log2(n))}: {ms:.3f} ms, {n/ms/1e6:.2f} G points/s, {launches} generated-kernel launches, '
      f'~{(fwd_bytes + bwd_bytes + 8 * n)/ms/1e6:.0f} GB/s incl. the torch sum; loss {loss.item():.3e} grad {b.grad.cpu().numpy()}')
