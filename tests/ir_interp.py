"""TEST INFRASTRUCTURE ONLY: a numpy interpreter for the straight-line IR (include/numga_b200.h).

Lets the CPU test-suite check what the tracer / lowering emit (against golden vectors from the
reference) without a GPU.  It is never imported by the product package: numga_b200 has no CPU
execution path.  Arithmetic is done op by op in the program dtype, i.e. the semantics of a
NGB200_FAITHFUL program; FAST programs differ only by FMA contraction.
"""
import numpy as np

from numga_b200 import runtime as rt


def run_ir(ir, inputs, dtype, params=()):
    """inputs: list of arrays [batch, width] (or [width] for broadcast operands). Returns list of [batch, width]."""
    t = np.dtype(dtype).type
    vals = []
    batch = max([np.atleast_2d(x).shape[0] for x in inputs] + [1])
    with np.errstate(all='ignore'):
        for op, a, b, c in ir['instr']:
            if op == rt.OP_INPUT:
                x = np.atleast_2d(np.asarray(inputs[a], dtype=dtype))
                v = np.broadcast_to(x[:, b], (batch,)).astype(dtype)
            elif op == rt.OP_CONST:
                v = np.full((batch,), t(ir['consts'][a]), dtype=dtype)
            elif op == rt.OP_PARAM:
                v = np.full((batch,), t(params[a]), dtype=dtype)
            elif op == rt.OP_ADD: v = vals[a] + vals[b]
            elif op == rt.OP_SUB: v = vals[a] - vals[b]
            elif op == rt.OP_MUL: v = vals[a] * vals[b]
            elif op == rt.OP_DIV: v = vals[a] / vals[b]
            elif op == rt.OP_NEG: v = -vals[a]
            elif op == rt.OP_SQRT: v = np.sqrt(vals[a])
            elif op == rt.OP_RSQRT: v = t(1) / np.sqrt(vals[a])
            elif op == rt.OP_RCP: v = t(1) / vals[a]
            elif op == rt.OP_ABS: v = np.abs(vals[a])
            elif op == rt.OP_NAN_TO_NUM:
                v = np.where(np.isnan(vals[a]), vals[b], np.nan_to_num(vals[a], nan=0.0))
            elif op == rt.OP_FMA: v = (vals[a].astype(np.longdouble) * vals[b] + vals[c]).astype(dtype)
            elif op == rt.OP_EXP: v = np.exp(vals[a])
            elif op == rt.OP_LOG: v = np.log(vals[a])
            elif op == rt.OP_SIN: v = np.sin(vals[a])
            elif op == rt.OP_COS: v = np.cos(vals[a])
            elif op == rt.OP_MIN: v = np.minimum(vals[a], vals[b])
            elif op == rt.OP_MAX: v = np.maximum(vals[a], vals[b])
            elif op == rt.OP_SELECT_GT0: v = np.where(vals[a] > 0, vals[b], vals[c])
            elif op == rt.OP_ATAN2: v = np.arctan2(vals[a], vals[b])
            elif op == rt.OP_PARTNER: v = vals[a].reshape(-1, 2)[:, ::-1].reshape(-1)      # items come in pairs (2k, 2k + 1)
            else: raise ValueError(op)
            vals.append(np.asarray(v, dtype=dtype))
    outs = [np.zeros((batch, w), dtype=dtype) for w in ir['out_width']]
    for k, c, v in ir['stores']:
        outs[k][:, c] = vals[v]
    return outs
