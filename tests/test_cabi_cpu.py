"""The C-ABI shared library: loads without a GPU, exports every symbol the header declares, compiles
kernels for sm_100a, and FAILS LOUDLY (no CPU fallback) when asked to launch without a device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from numga_b200 import runtime as rt
from numga_b200.algebra import Algebra

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'numga_b200.h')


def _declared():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ngb200_[a-z_0-9]+)\s*\(', text)))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(rt.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/numga_b200.h but not exported'
    assert set(rt.EXPORTED) == set(declared), 'runtime.py binds a different set of symbols than the header declares'


def test_library_contains_sm100a_code_and_no_driver_link_dependency():
    out = subprocess.run(['cuobjdump', '-lelf', rt.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    needed = subprocess.run(['readelf', '-d', rt.LIB_PATH], capture_output=True, text=True).stdout
    assert 'libcuda.so' not in needed      # the driver is dlopen()ed at first launch, so this loads on CPU boxes


def test_version_string():
    assert rt.version().startswith('numga_b200 ') and 'sm_100a' in rt.version()


def _terms(op):
    return np.concatenate([op.index, op.coeff[:, None]], axis=1)


def test_operator_kernel_compiles_without_gpu_and_is_vectorised():
    a = Algebra.from_pqr(3, 0, 1)
    M = a.subspace.even_grade()
    prog = rt.Program.from_terms(_terms(a.operator.product(M, M)), [8, 8], 8, rt.F32)
    info = prog.info
    assert info['vec'] == 4 and info['n_instr'] == 88 and info['cubin_bytes'] > 0
    src = prog.source
    assert 'float4' in src and 'ngb_vec' in src and 'ngb_scalar' in src


def test_generated_sass_uses_128bit_global_accesses(tmp_path):
    a = Algebra.from_pqr(3, 0, 1)
    M = a.subspace.even_grade()
    os.environ['NGB200_CACHE_DIR'] = str(tmp_path)
    try:
        rt.check(rt.lib.ngb200_set_cache_dir(str(tmp_path).encode()))
        rt.Program.from_terms(_terms(a.operator.product(M, M)), [8, 8], 8, rt.F32)
        cubins = [f for f in os.listdir(tmp_path) if f.endswith('.cubin')]
        assert len(cubins) == 1
        sass = subprocess.run(['cuobjdump', '-sass', os.path.join(tmp_path, cubins[0])], capture_output=True, text=True).stdout
        assert sass.count('LDG.E.128') >= 16 and sass.count('STG.E.128') >= 8
    finally:
        del os.environ['NGB200_CACHE_DIR']
        rt.check(rt.lib.ngb200_set_cache_dir(b''))


def test_kernel_variants_per_program_class():
    """Which kernels the generator emits ('Kernels'): small programs a 128-bit vector kernel; wide programs
    whose two items do not fit in 128 registers a ONE-item contiguous kernel (batch stride compiled in, blockDim.x
    index arithmetic); gathered programs a second scalar kernel with the batch stride compiled in."""
    pga, cga = Algebra.from_pqr(3, 0, 1), Algebra.from_pqr(4, 1, 0)
    M, F = pga.subspace.even_grade(), cga.subspace.full()
    small = rt.Program.from_terms(_terms(pga.operator.product(M, M)), [8, 8], 8, rt.F32)
    assert small.info['vec'] == 4 and 'float4' in small.source and 'ngb_scalar_u' not in small.source
    wide = rt.Program.from_terms(_terms(cga.operator.product(F, F)), [32, 32], 32, rt.F32)
    src = wide.source
    assert wide.info['vec'] == 1 and 'ngb_vec(const Params p)' in src and '#define NGB_BDIM blockDim.x' in src
    assert 'x[0] = NGB_LD(p.in[0].ptr + 0LL * p.in[0].cs + v);' in src            # no run-time batch stride
    assert 'ngb_scalar(const Params p)' in src                                    # strided fallback still there
    # gathered: out[i] = a[index[i]] + b[i]
    instr = [(rt.OP_INPUT, 0, 0, 0), (rt.OP_INPUT, 1, 0, 0), (rt.OP_ADD, 0, 1, 0)]
    gathered = rt.Program.from_ir(rt.F32, rt.FAST, [1, 1], [rt.INDEXED, rt.VARYING], [1], [rt.VARYING], 0, [], instr, [(0, 0, 2)])
    src = gathered.source
    assert 'ngb_scalar_u(const Params p)' in src and 'ngb_vec' not in src
    generic, unit = src.split('ngb_scalar_u(const Params p)')
    assert 'p.in[0].index[i] NGB_BS(p.in[0].bs)' in generic and '#define NGB_BS(b) * (b)' in generic
    assert unit.count('#define NGB_BS') == 0 and '#define NGB_BS(b) \n' in generic.split('ngb_scalar(const Params p)')[1]


def test_bad_descriptions_are_rejected_with_a_message():
    with pytest.raises(rt.NGB200Error, match='out of range'):
        rt.Program.from_terms([[9, 0, 0, 1]], [8, 8], 8, rt.F32)
    with pytest.raises(rt.NGB200Error, match='operand'):
        rt.Program.from_ir(rt.F32, rt.FAST, [2], [0], [1], [0], 0, [], [(rt.OP_INPUT, 0, 0, 0), (rt.OP_ADD, 0, 5, 0)], [(0, 0, 1)])
    with pytest.raises(rt.NGB200Error, match='dtype'):
        rt.Program.from_ir(7, rt.FAST, [2], [0], [1], [0], 0, [], [(rt.OP_INPUT, 0, 0, 0)], [(0, 0, 0)])


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_launch_without_gpu_fails_loudly():
    a = Algebra.from_pqr(3, 0, 1)
    M = a.subspace.even_grade()
    prog = rt.Program.from_terms(_terms(a.operator.product(M, M)), [8, 8], 8, rt.F32)
    x = np.zeros((8, 16), np.float32)
    with pytest.raises(rt.NGB200Error, match='no CPU fallback'):
        prog.launch([(x.ctypes.data, 16, 1)] * 2, [(x.ctypes.data, 16, 1)], 16)
    from numga_b200 import B200Context
    ctx = B200Context((3, 0, 1), device='cpu')
    m = ctx.multivector.motor(np.ones((4, 8), np.float32))
    with pytest.raises(rt.NGB200Error, match='no CPU path'):
        m * m


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'numga_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.replace('numga_oracle', 'oracle') or f == 'ngb200.cu' and 'oracle' not in text, \
                    f'{f} mentions the oracle: product code must not depend on it'


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The header is plain C (compiles with gcc, no C++ / CUDA types) and the ctypes mirror in runtime.py has the same
    struct sizes and field offsets."""
    import ctypes
    import subprocess
    src = tmp_path / 'layout.c'
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "numga_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ngb200_operand), offsetof(ngb200_operand, comp_stride),
           offsetof(ngb200_operand, batch_stride), offsetof(ngb200_operand, index), offsetof(ngb200_operand, inner_size),
           offsetof(ngb200_operand, outer_stride), sizeof(ngb200_instr));
    printf("%zu %zu %zu %zu\\n", sizeof(ngb200_program_desc), offsetof(ngb200_program_desc, consts),
           offsetof(ngb200_program_desc, stores), offsetof(ngb200_program_desc, vec));
    printf("%zu %zu\\n", sizeof(ngb200_program_info), offsetof(ngb200_program_info, cubin_bytes));
    printf("%d %d %d %d %d\\n", NGB200_VARYING, NGB200_BROADCAST, NGB200_INDEXED, NGB200_REDUCED, NGB200_TILED);
    printf("%d %d %d %d %d %d\\n", NGB200_OP_INPUT, NGB200_OP_FMA, NGB200_OP_SELECT_GT0, NGB200_OP_ATAN2, NGB200_OP_PARTNER, NGB200_OP_COUNT_);
    return 0;
}
''')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split('\n')
    op = rt.Operand
    assert lines[0].split() == [str(x) for x in (ctypes.sizeof(op), op.comp_stride.offset, op.batch_stride.offset, op.index.offset,
                                                  op.inner_size.offset, op.outer_stride.offset, ctypes.sizeof(rt.Instr))]
    d = rt.ProgramDesc
    assert lines[1].split() == [str(x) for x in (ctypes.sizeof(d), d.consts.offset, d.stores.offset, d.vec.offset)]
    assert lines[2].split() == [str(ctypes.sizeof(rt.ProgramInfo)), str(rt.ProgramInfo.cubin_bytes.offset)]
    assert lines[3].split() == [str(x) for x in (rt.VARYING, rt.BROADCAST, rt.INDEXED, rt.REDUCED, rt.TILED)]
    assert lines[4].split() == [str(x) for x in (rt.OP_INPUT, rt.OP_FMA, rt.OP_SELECT_GT0, rt.OP_ATAN2, rt.OP_PARTNER, rt.OP_PARTNER + 1)]


def test_every_opcode_generates_code_in_both_precisions():
    """One program per arithmetic opcode of include/numga_b200.h (compile only, no GPU): the generator knows them all,
    in the scalar and in the packed (two items per thread) emitters."""
    ops = [getattr(rt, n) for n in dir(rt) if n.startswith('OP_') and n not in ('OP_INPUT', 'OP_CONST', 'OP_PARAM', 'OP_PARTNER')]
    assert len(ops) == rt.OP_PARTNER - 3           # (OP_PARTNER exists in pipeline phases only: tested there)
    for dtype in (rt.F32, rt.F64):
        for op in ops:
            n = rt.N_OPERANDS.get(op, 2)
            instr = [(rt.OP_INPUT, 0, k, 0) for k in range(3)] + [(op, 0, 1 if n > 1 else 0, 2 if n > 2 else 0)]
            mode = rt.FAST if op == rt.OP_FMA else rt.FAITHFUL
            for m in {mode, rt.FAST}:
                prog = rt.Program.from_ir(dtype, m, [3], [rt.VARYING], [1], [rt.VARYING], 0, [], instr, [(0, 0, 3)])
                assert prog.info['cubin_bytes'] > 0, (op, dtype, m)


def test_pipeline_c_abi_compiles_and_validates():
    """ngb200_pipeline_*: two phases over two domains exchanging through a shared buffer compile into ONE sm_100a
    kernel (barrier between the phases, gather through an index operand); malformed descriptions are refused."""
    I, S = rt.BIND_GLOBAL, rt.BIND_SHARED
    # phase 0 (domain 0, "nodes"): s = 2 * a                       -> shared buffer 0
    # phase 1 (domain 1, "edges"): out = s[i0] - s[i1] (gathered)  -> global
    scale = dict(in_width=[1], out_width=[1], n_params=0, consts=[2.0], stores=[(0, 0, 2)],
                 instr=[(rt.OP_INPUT, 0, 0, 0), (rt.OP_CONST, 0, 0, 0), (rt.OP_MUL, 0, 1, 0)])
    diff = dict(in_width=[1, 1], out_width=[1], n_params=0, consts=[], stores=[(0, 0, 2)],
                instr=[(rt.OP_INPUT, 0, 0, 0), (rt.OP_INPUT, 1, 0, 0), (rt.OP_SUB, 0, 1, 0)])
    globals_ = [(1, 0, False), (1, 1, True), (1, 1, True), (1, 1, False)]       # a, i0, i1, out
    phases = [(scale, 0, [(I, 0, -1)], [(S, 0, -1)]), (diff, 1, [(S, 0, 1), (S, 0, 2)], [(I, 3, -1)])]
    p = rt.PipelineProgram(rt.F32, rt.FAST, [64, 63], globals_, [(1, 0)], phases, (0, 0), 0)
    src = p.source
    assert src.count('__syncthreads()') == 2 and 'ngb_pipeline' in src and 'p.idx[0][gi] - tile * 64LL' in src
    info = p.info
    # record layout: one 16-byte chunk per item (odd chunk count) + one padding chunk per 8 rows + 1
    assert info['shared_bytes'] == (64 * 1 + 64 // 8 + 1) * 16 and info['threads_per_tile'] == 64 and info['n_instr'] >= 2
    assert 'float4' in src                      # 128-bit shared accesses
    with pytest.raises(rt.NGB200Error, match='no CPU fallback|CUDA'):
        p.launch([(1, 1, 1)] * 4, [64, 63], 1)
    # tiles that need only one warp: several tiles per CTA, each warp with its own shared-memory slice, warp barriers
    w = rt.PipelineProgram(rt.F32, rt.FAST, [32, 31], globals_, [(1, 0)], phases, (0, 0), 0)
    assert '__syncwarp()' in w.source and '__syncthreads()' not in w.source and '(threadIdx.x >> 5)' in w.source
    assert w.info['threads_per_tile'] == 32 and w.info['shared_bytes'] == 8 * ((32 + 4 + 1) * 16)
    bad = [(scale, 0, [(I, 0, -1)], [(S, 0, -1)]), (diff, 1, [(S, 0, -1), (S, 0, 2)], [(I, 3, -1)])]
    with pytest.raises(rt.NGB200Error, match='another domain'):
        rt.PipelineProgram(rt.F32, rt.FAST, [64, 63], globals_, [(1, 0)], bad, (0, 0), 0)
    with pytest.raises(rt.NGB200Error, match='loop'):
        rt.PipelineProgram(rt.F32, rt.FAST, [64, 63], globals_, [(1, 0)], phases, (1, 3), 0)
    with pytest.raises(rt.NGB200Error, match='shared memory'):
        rt.PipelineProgram(rt.F32, rt.FAST, [1024, 63], globals_, [(60, 0)], phases, (0, 0), 0)


def test_partner_opcode_is_a_warp_convergent_pipeline_feature():
    """NGB200_OP_PARTNER (value of the paired item i ^ 1, one warp shuffle): accepted in pipeline phases, where the phase
    loop is emitted so that EVERY lane executes the body (lanes without an item compute on the tile's first item and
    store nothing); refused in ordinary programs, whose kernels map several items to a thread."""
    I = rt.BIND_GLOBAL
    swap = dict(in_width=[1], out_width=[1], n_params=0, consts=[], stores=[(0, 0, 2)],
                instr=[(rt.OP_INPUT, 0, 0, 0), (rt.OP_PARTNER, 0, 0, 0), (rt.OP_SUB, 0, 1, 0)])
    p = rt.PipelineProgram(rt.F64, rt.FAITHFUL, [70], [(1, 0, False), (1, 0, False)], [], [(swap, 0, [(I, 0, -1)], [(I, 1, -1)])], (0, 0), 0)
    src = p.source
    assert '__shfl_xor_sync(0xffffffffu' in src and 'ngb_on' in src and 'if (ngb_on) {' in src
    assert 'for (int ngb_i0 = 0; ngb_i0 < 70; ngb_i0 += NGB_TPB)' in src          # uniform trip count for all lanes
    with pytest.raises(rt.NGB200Error, match='pipeline phases'):
        rt.Program.from_ir(rt.F32, rt.FAST, [1], [rt.VARYING], [1], [rt.VARYING], 0, [], swap['instr'], swap['stores'])
