"""Generate the committed golden fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py

Writes, next to this file:
  tables.npz        integer tables (axes blades + non-zero terms) of operators built by the reference's
                    OperatorFactory, for the algebras / subspace tuples the hot path touches and more;
  ref_fixtures.npz  the reference's OWN pinned operator tables (operators/<pqr>/*.txt, written by
                    numga/operator/test/test_pinning.py), parsed into (left blade, right blade) ->
                    (sign, output blade) arrays;
  vectors.npz       seeded float inputs and the reference numpy backend's outputs (fp32 and fp64)
                    for every case in tests/cases.py;
  MANIFEST.json     what was generated, numpy version, reference path.
The GPU box has no /root/reference: tests only read these files.
"""
import glob
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ref_loader import REFERENCE, load_reference  # noqa: E402

numga = load_reference()
from numga.algebra.algebra import Algebra  # noqa: E402
from numga.backend.numpy.context import NumpyContext  # noqa: E402

from cases import CASES, case_inputs  # noqa: E402

TABLE_ALGEBRAS = [(2, 0, 0), (3, 0, 0), (2, 0, 1), (3, 0, 1), (1, 1, 1), (3, 1, 0), (1, 1, 0), (4, 1, 0)]
SUBSPACES = ['scalar', 'vector', 'bivector', 'pseudoscalar', 'antivector', 'even_grade', 'odd_grade', 'multivector',
             'scalar_pseudoscalar', 'bireflection', 'self_reverse']
UNARY = ['reverse', 'involute', 'conjugate', 'right_hodge', 'left_hodge', 'right_hodge_inverse', 'right_complement',
         'left_complement', 'scalar_negation', 'pseudoscalar_negation', 'degenerate', 'nondegenerate', 'squared',
         'symmetric_reverse_product', 'symmetric_scalar_negation_product', 'symmetric_conjugate_product',
         'study_norm_squared', 'inverse_factor']
BINARY = ['geometric_product', 'wedge_product', 'inner_product', 'scalar_product', 'regressive_product',
          'reverse_product', 'left_contraction_product', 'right_contraction_product', 'anti_geometric_product',
          'anti_wedge_product', 'commutator_product', 'anti_commutator_product', 'bivector_product', 'cross_product',
          'scalar_negation_product', 'sandwich', 'reverse_sandwich', 'full_sandwich', 'inertia', 'select', 'restrict']


def pack_operator(op):
    k = np.asarray(op.kernel).astype(np.int64)
    nz = np.nonzero(k)
    return dict(
        axes=np.concatenate([a.blades.astype(np.int64) for a in op.axes]) if op.axes else np.zeros(0, np.int64),
        axes_len=np.array([len(a) for a in op.axes], np.int64),
        index=np.stack(nz, axis=1).astype(np.int16).reshape(-1, k.ndim),
        coeff=k[nz].astype(np.int16))


def make_tables():
    """All tables go into a handful of flat arrays (an .npz member per table costs more than the table)."""
    keys, arity, axes_len, axes, nnz, index, coeff = [], [], [], [], [], [], []

    def add(key, op):
        p = pack_operator(op)
        keys.append(key)
        arity.append(len(op.axes) - 1)
        axes_len.append(p['axes_len'])
        axes.append(p['axes'])
        nnz.append(len(p['coeff']))
        index.append(p['index'].reshape(-1))
        coeff.append(p['coeff'])

    for pqr in TABLE_ALGEBRAS:
        alg = Algebra.from_pqr(*pqr)
        tag = ''.join(map(str, pqr))
        big = sum(pqr) >= 5
        subs = {s: getattr(alg.subspace, s)() for s in SUBSPACES}
        for name in UNARY:
            for s, sub in subs.items():
                if big and s in ('multivector', 'odd_grade') and name == 'inverse_factor':
                    continue
                try:
                    op = getattr(alg.operator, name)(sub)
                except Exception:
                    continue
                add(f'{tag}/{name}/{s}', op)
        for name in BINARY:
            for (s1, a), (s2, b) in itertools.product(subs.items(), subs.items()):
                if big and 'multivector' in (s1, s2) and name in ('sandwich', 'reverse_sandwich', 'full_sandwich', 'inertia'):
                    continue
                try:
                    op = getattr(alg.operator, name)(a, b)
                except Exception:
                    continue
                add(f'{tag}/{name}/{s1},{s2}', op)
    np.savez_compressed(
        os.path.join(HERE, 'tables.npz'), keys=np.array(keys), arity=np.array(arity, np.int8),
        axes_len=np.concatenate(axes_len).astype(np.int16), axes=np.concatenate(axes).astype(np.int16),
        nnz=np.array(nnz, np.int32), index=np.concatenate(index).astype(np.int8),
        coeff=np.concatenate(coeff).astype(np.int8))
    return len(keys)


def parse_fixture(path):
    """operators/<pqr>/<name>.txt -> (left blades, right blades, sign[l, r], out blade[l, r]); the format is
    the pretty-printer of numga/operator/operator.py:246-331 (blade strings are LSB-first bit strings)."""
    lines = open(path, encoding='utf-8').read().split('\n')
    rows = [[c for c in line.replace('│', '║').split('║')[1:-1]] for line in lines[1::2]]
    blade = lambda s: int(s.strip()[::-1], 2)
    right = [blade(c) for c in rows[0][1:]]
    body = rows[1:]
    unary = len(body) == 1 and body[0][0].strip() == ''
    left = [0] if unary else [blade(r[0]) for r in body]
    sign = np.zeros((len(left), len(right)), np.int8)
    out = np.zeros((len(left), len(right)), np.int64)
    for i, r in enumerate(body):
        for j, cell in enumerate(r[1:]):
            cell = cell.strip()
            if cell:
                sign[i, j] = {'+': 1, '-': -1}[cell[0]]
                out[i, j] = blade(cell[1:])
    return np.array(left, np.int64), np.array(right, np.int64), sign, out, unary


def make_ref_fixtures():
    out, count = {}, 0
    for path in sorted(glob.glob(os.path.join(REFERENCE, 'operators', '*', '*.txt'))):
        pqr, name = path.split(os.sep)[-2], os.path.basename(path)[:-4]
        left, right, sign, blades, unary = parse_fixture(path)
        out[f'{pqr}/{name}/left'] = left
        out[f'{pqr}/{name}/right'] = right
        out[f'{pqr}/{name}/sign'] = sign
        out[f'{pqr}/{name}/blade'] = blades
        out[f'{pqr}/{name}/unary'] = np.array(int(unary))
        count += 1
    np.savez_compressed(os.path.join(HERE, 'ref_fixtures.npz'), **out)
    return count


def make_vectors():
    out = {}
    for case in CASES:
        name, pqr, specs, scale, fn = case
        for dtype in (np.float32, np.float64):
            ctx = NumpyContext(pqr, dtype=dtype)
            subs = [getattr(ctx.subspace, s)() for s, _ in specs]
            arrays = case_inputs(case, dtype)([len(s) for s in subs])
            mvs = [ctx.multivector(values=a, subspace=s) for a, s in zip(arrays, subs)]
            with np.errstate(all='ignore'):
                res = fn(*mvs)
            tag = f'{name}/{np.dtype(dtype).name}'
            for k, a in enumerate(arrays):
                out[f'{tag}/in{k}'] = a
            out[f'{tag}/out'] = np.ascontiguousarray(np.broadcast_to(res.values, res.values.shape))
            assert res.values.dtype == dtype, (name, res.values.dtype)
            out[f'{tag}/out_blades'] = res.subspace.blades.astype(np.int64)
    np.savez_compressed(os.path.join(HERE, 'vectors.npz'), **out)
    return len(CASES)


if __name__ == '__main__':
    manifest = dict(reference=REFERENCE, numpy=np.__version__, tables=make_tables(), ref_fixtures=make_ref_fixtures(),
                    vector_cases=make_vectors(),
                    note='generated by tests/golden/make_golden.py from the unmodified reference (patched import only)')
    json.dump(manifest, open(os.path.join(HERE, 'MANIFEST.json'), 'w'), indent=1)
    print(manifest)
