"""Readers for the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os
from functools import lru_cache

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@lru_cache(None)
def tables():
    """{key: (axes [list of blade arrays], index [nnz, arity+1], coeff [nnz])}; key = 'pqr/op/sub[,sub]'."""
    z = np.load(os.path.join(HERE, 'tables.npz'))
    keys, arity, axes_len, axes, nnz, index, coeff = (z[k] for k in ('keys', 'arity', 'axes_len', 'axes', 'nnz', 'index', 'coeff'))
    out, pl, pa, pi, pc = {}, 0, 0, 0, 0
    for key, ar, n in zip(keys, arity.tolist(), nnz.tolist()):
        lens = axes_len[pl:pl + ar + 1]; pl += ar + 1
        ax = []
        for ln in lens.tolist():
            ax.append(axes[pa:pa + ln].astype(np.int64)); pa += ln
        idx = index[pi:pi + n * (ar + 1)].reshape(n, ar + 1).astype(np.int64); pi += n * (ar + 1)
        out[str(key)] = (ax, idx, coeff[pc:pc + n].astype(np.int64)); pc += n
    return out


@lru_cache(None)
def ref_fixtures():
    z = np.load(os.path.join(HERE, 'ref_fixtures.npz'))
    names = sorted({k.rsplit('/', 1)[0] for k in z.files})
    return {n: {f: z[f'{n}/{f}'] for f in ('left', 'right', 'sign', 'blade', 'unary')} for n in names}


@lru_cache(None)
def vectors():
    return np.load(os.path.join(HERE, 'vectors.npz'))


def case_vectors(name, dtype):
    z = vectors()
    tag = f'{name}/{np.dtype(dtype).name}'
    ins = []
    while f'{tag}/in{len(ins)}' in z.files:
        ins.append(z[f'{tag}/in{len(ins)}'])
    return ins, z[f'{tag}/out'], z[f'{tag}/out_blades']


def dense_of(axes, index, coeff):
    k = np.zeros([len(a) for a in axes], dtype=np.int64)
    if len(coeff):
        k[tuple(index.T)] = coeff
    return k
