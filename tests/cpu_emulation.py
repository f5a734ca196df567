"""TEST INFRASTRUCTURE ONLY: run the product's host logic on CPU tensors by replacing kernel launches with the
numpy IR interpreter (tests/ir_interp.py).

The product (numga_b200/) has no CPU execution path and no hook for one; this module monkeypatches
`CompiledFunction.launch` / `B200Operator.__call__` from the outside, inside a context manager, so that the CPU
test tier can check orchestration logic numerically (broadcasting, gather/scatter, the physics step) against golden
data without a GPU.  The arithmetic is that of the IR, op by op in the context dtype (= FAITHFUL semantics).
"""
import contextlib
import os

import numpy as np
import torch

from ir_interp import run_ir
from numga_b200 import backend as B
from numga_b200 import runtime as rt


@contextlib.contextmanager
def emulate_launches():
    orig_call, orig_op_call = B.CompiledFunction.launch, B.B200Operator.__call__

    def compiled_call(self, args, bshape=None, out=None):
        ctx = self.context
        mvs = [(role[1], a) for role, a in zip(self.arg_roles, args) if role[0] == 'mv']
        params = [float(a) for role, a in zip(self.arg_roles, args) if role[0] == 'param']
        if bshape is None:
            bshape = B._broadcast([B._batch_shape(a) for _, a in mvs])
        n_batch = B._prod(bshape)
        np_dtype = B._NP_DTYPE[ctx.dtype]
        inputs = [None] * len(mvs)
        for k, a in mvs:
            if isinstance(a, B.Indexed):                       # gather
                g = a.as_mv().values.detach().cpu().numpy()[a.index.cpu().numpy()]
                inputs[k] = g.reshape(-1, g.shape[-1])
                continue
            v = a.values.detach().cpu().numpy()
            if self.program.in_mode[k] == rt.BROADCAST:
                inputs[k] = v.reshape(-1)
            else:
                inputs[k] = np.broadcast_to(v, tuple(bshape) + (v.shape[-1],)).reshape(n_batch, v.shape[-1])
        outs = run_ir(self.ir, inputs, np_dtype, params) if n_batch else [np.zeros((0, w), np_dtype) for w in self.ir['out_width']]
        for k, mode in enumerate(self.program.out_mode):      # NGB200_REDUCED outputs: the batch sum
            if mode == rt.REDUCED:
                outs[k] = outs[k][:max(n_batch, 0)].sum(axis=0, keepdims=True)
        result = []
        if out is not None:                                    # in-place destinations (scatter for Indexed)
            for dest, o in zip(out, outs):
                if isinstance(dest, B.Indexed):
                    rows = torch.as_tensor(np.ascontiguousarray(o)).to(ctx.dtype)
                    dest.as_mv().values[dest.index] = rows.reshape(tuple(dest.index.shape) + (rows.shape[-1],))
                    result.append(dest.target)
                else:
                    dest.values.copy_(torch.as_tensor(np.ascontiguousarray(o)).reshape(dest.values.shape))
                    result.append(dest)
            return result[0] if self.single else tuple(result)
        for k, (s, o) in enumerate(zip(self.out_subspaces, outs)):
            if self.program.out_mode[k] == rt.REDUCED:
                result.append(ctx.mvtype(ctx, s, torch.as_tensor(np.ascontiguousarray(o[0])).to(ctx.dtype)))
                continue
            t = B.to_soa(torch.as_tensor(np.ascontiguousarray(o[:max(n_batch, 1)])).reshape(tuple(bshape) + (len(s),))) \
                if n_batch > 1 else torch.as_tensor(np.ascontiguousarray(o[:n_batch].reshape(tuple(bshape) + (len(s),))))
            result.append(ctx.mvtype(ctx, s, t.to(ctx.dtype)))
        return result[0] if self.single else tuple(result)

    def operator_call(self, *inputs):
        if any(isinstance(i, (B.SubSpace, B.TracedMultiVector)) for i in inputs) or len(self.output) == 0:
            return orig_op_call(self, *inputs)
        return self.context.engine.call(('operator', self.operator.uid), lambda *xs: orig_op_call(self, *xs), inputs)

    def pipeline_launch(self, arrays, items, repeat=1, params=()):
        """Pipelines (numga_b200/pipeline.py): tiles are independent and phases are separated by barriers, so running
        every phase over the WHOLE domain with shared buffers as full-size arrays is the same computation."""
        pipe, ctx = self.pipe, self.context
        np_dtype = B._NP_DTYPE[ctx.dtype]
        counts = [int(items[d]) for d in pipe.domains]
        G = []
        for name, width, domain, is_index in pipe.globals:
            a = arrays[name]
            G.append(a if is_index else (a.flat if isinstance(a, B.B200DenseOperator) else a.values))
        S = [np.zeros((counts[d], w), np_dtype) for w, d in pipe.shared]

        def read(bind, n):
            kind, slot, index = bind
            src = S[slot] if kind == rt.BIND_SHARED else G[slot].detach().cpu().numpy()
            if kind == rt.BIND_GLOBAL and pipe.globals[slot][2] < 0:
                return src.reshape(-1)
            src = src.reshape(-1, src.shape[-1])
            return src[G[index].cpu().numpy()[:n]] if index >= 0 else src[:n]

        def write(bind, n, values):
            kind, slot, index = bind
            if kind == rt.BIND_SHARED:
                if index >= 0:
                    S[slot][G[index].cpu().numpy()[:n]] = values
                else:
                    S[slot][:n] = values
            else:
                rows = torch.as_tensor(np.ascontiguousarray(values)).to(ctx.dtype)
                if index >= 0:
                    G[slot][G[index][:n]] = rows
                else:
                    G[slot][:n] = rows

        def run(k):
            ir, d, ins, outs = self.phases[k]
            n = counts[d]
            res = run_ir(ir, [read(b, n) for b in ins], np_dtype, [float(p) for p in params])
            for b, o in zip(outs, res):
                write(b, n, o[:n])
        lo, hi = self.loop
        for k in range(lo):
            run(k)
        for _ in range(repeat):
            for k in range(lo, hi):
                run(k)
        for k in range(hi, len(self.phases)):
            run(k)
        return []

    from numga_b200 import pipeline as PL
    orig_pipe = PL.CompiledPipeline.launch
    B.CompiledFunction.launch, B.B200Operator.__call__, PL.CompiledPipeline.launch = compiled_call, operator_call, pipeline_launch
    try:
        yield
    finally:
        B.CompiledFunction.launch, B.B200Operator.__call__, PL.CompiledPipeline.launch = orig_call, orig_op_call, orig_pipe


def emulated_context(algebra, dtype=torch.float64, mode='faithful'):
    """A context on CPU tensors for use inside `emulate_launches()`."""
    ctx = B.B200Context(algebra, dtype=dtype, device='cpu', mode=mode, compile_only=True)
    ctx.compile_only = False     # every launch site is patched; host-side torch set-up work (e.g. inverse()) runs for real
    ctx.lazy = os.environ.get('NGB200_TEST_LAZY', '0') == '1'      # rerun any CPU-tier test under automatic fusion
    return ctx
