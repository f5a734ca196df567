"""ctypes binding of the C ABI in include/numga_b200.h (libnumga_b200.so, built in-tree).

This is the only way Python reaches the GPU kernels.  There is no fallback: if the shared
library is missing, importing this module raises; if there is no CUDA device, launching raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libnumga_b200.so')

F32, F64 = 0, 1
FAST, FAITHFUL = 0, 1
VARYING, BROADCAST, INDEXED, REDUCED, TILED = 0, 1, 2, 3, 4

(OP_INPUT, OP_CONST, OP_PARAM, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SQRT, OP_RSQRT, OP_RCP, OP_ABS,
 OP_NAN_TO_NUM, OP_FMA, OP_EXP, OP_LOG, OP_SIN, OP_COS, OP_MIN, OP_MAX, OP_SELECT_GT0, OP_ATAN2, OP_PARTNER) = range(23)

N_OPERANDS = {OP_INPUT: 0, OP_CONST: 0, OP_PARAM: 0, OP_NEG: 1, OP_SQRT: 1, OP_RSQRT: 1, OP_RCP: 1, OP_ABS: 1,
              OP_EXP: 1, OP_LOG: 1, OP_SIN: 1, OP_COS: 1, OP_FMA: 3, OP_SELECT_GT0: 3, OP_PARTNER: 1}


class NGB200Error(RuntimeError):
    pass


class Instr(C.Structure):
    _fields_ = [('op', C.c_int32), ('a', C.c_int32), ('b', C.c_int32), ('c', C.c_int32)]


class Store(C.Structure):
    _fields_ = [('operand', C.c_int32), ('component', C.c_int32), ('value', C.c_int32)]


class ProgramDesc(C.Structure):
    _fields_ = [
        ('dtype', C.c_int32), ('flags', C.c_uint32),
        ('n_inputs', C.c_int32), ('input_width', C.POINTER(C.c_int32)), ('input_mode', C.POINTER(C.c_int32)),
        ('n_outputs', C.c_int32), ('output_width', C.POINTER(C.c_int32)), ('output_mode', C.POINTER(C.c_int32)),
        ('n_params', C.c_int32),
        ('n_consts', C.c_int32), ('consts', C.POINTER(C.c_double)),
        ('n_instr', C.c_int32), ('instr', C.POINTER(Instr)),
        ('n_stores', C.c_int32), ('stores', C.POINTER(Store)),
        ('vec', C.c_int32),
    ]


class Operand(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('comp_stride', C.c_int64), ('batch_stride', C.c_int64), ('index', C.c_void_p),
                ('inner_size', C.c_int64), ('outer_stride', C.c_int64)]


class Bind(C.Structure):
    _fields_ = [('kind', C.c_int32), ('slot', C.c_int32), ('index', C.c_int32)]


class PhaseDesc(C.Structure):
    _fields_ = [('program', ProgramDesc), ('domain', C.c_int32), ('inputs', C.POINTER(Bind)), ('outputs', C.POINTER(Bind))]


class PipelineDesc(C.Structure):
    _fields_ = [
        ('dtype', C.c_int32), ('flags', C.c_uint32),
        ('n_domains', C.c_int32), ('items_per_tile', C.POINTER(C.c_int32)),
        ('n_globals', C.c_int32), ('global_width', C.POINTER(C.c_int32)), ('global_domain', C.POINTER(C.c_int32)),
        ('global_is_index', C.POINTER(C.c_int32)),
        ('n_shared', C.c_int32), ('shared_width', C.POINTER(C.c_int32)), ('shared_domain', C.POINTER(C.c_int32)),
        ('n_phases', C.c_int32), ('phases', C.POINTER(PhaseDesc)),
        ('loop_begin', C.c_int32), ('loop_end', C.c_int32), ('n_params', C.c_int32), ('threads_per_tile', C.c_int32),
    ]


class PipelineInfo(C.Structure):
    _fields_ = [('n_instr', C.c_int32), ('threads_per_tile', C.c_int32), ('min_blocks', C.c_int32), ('regs', C.c_int32),
                ('grid', C.c_int32), ('cache_hit', C.c_int32), ('shared_bytes', C.c_int64), ('spill_bytes', C.c_int64),
                ('local_bytes', C.c_int64)]


class ProgramInfo(C.Structure):
    _fields_ = [('n_instr', C.c_int32), ('n_uniform', C.c_int32), ('vec', C.c_int32), ('regs_vec', C.c_int32),
                ('regs_scalar', C.c_int32), ('cache_hit', C.c_int32), ('cubin_bytes', C.c_int64)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise NGB200Error(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(numga_b200 has no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
    sigs = {
        'ngb200_version': (C.c_char_p, []),
        'ngb200_last_error': (C.c_char_p, []),
        'ngb200_set_cache_dir': (i32, [C.c_char_p]),
        'ngb200_operator_create': (i32, [C.POINTER(i32), i32, i32, C.POINTER(i32), i32, i32, u32, u32, C.POINTER(vp)]),
        'ngb200_program_create': (i32, [C.POINTER(ProgramDesc), C.POINTER(vp)]),
        'ngb200_program_destroy': (None, [vp]),
        'ngb200_program_get_info': (i32, [vp, C.POINTER(ProgramInfo)]),
        'ngb200_program_source': (C.c_char_p, [vp]),
        'ngb200_program_launch': (i32, [vp, C.POINTER(Operand), C.POINTER(Operand), C.POINTER(C.c_double), i64, vp]),
        'ngb200_program_run_host': (i32, [vp, C.POINTER(Operand), C.POINTER(Operand), C.POINTER(C.c_double), i64, i32, i64]),
        'ngb200_host_copy_control': (i32, [vp, i64, vp, i64, i64, i32, i32]),
        'ngb200_plan_create': (i32, [vp, C.POINTER(Operand), C.POINTER(Operand), i64, C.POINTER(vp)]),
        'ngb200_plan_launch': (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_double), vp]),
        'ngb200_plan_destroy': (None, [vp]),
        'ngb200_pipeline_create': (i32, [C.POINTER(PipelineDesc), C.POINTER(vp)]),
        'ngb200_pipeline_destroy': (None, [vp]),
        'ngb200_pipeline_get_info': (i32, [vp, C.POINTER(PipelineInfo)]),
        'ngb200_pipeline_source': (C.c_char_p, [vp]),
        'ngb200_pipeline_launch': (i32, [vp, C.POINTER(Operand), C.POINTER(i64), i64, i32, C.POINTER(C.c_double), vp]),
        'ngb200_host_alloc': (i32, [C.POINTER(vp), i64]),
        'ngb200_host_free': (i32, [vp]),
        'ngb200_fill_normal': (i32, [vp, i64, u64, u64, C.c_double, i32, vp]),
        'ngb200_launch_count': (i64, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib, tuple(sigs)


lib, EXPORTED = _load()


def check(rc):
    if rc != 0:
        raise NGB200Error(f'numga_b200 error {rc}: {lib.ngb200_last_error().decode()}')


def version():
    return lib.ngb200_version().decode()


def launch_count():
    return int(lib.ngb200_launch_count())


def _i32(seq):
    seq = list(seq)
    return (C.c_int32 * max(len(seq), 1))(*seq)


BIND_GLOBAL, BIND_SHARED = 0, 1


def _program_desc(dtype, flags, in_width, in_mode, out_width, out_mode, n_params, consts, instr, stores, vec, keep):
    """ProgramDesc over ctypes arrays; the arrays are appended to `keep` (they must outlive the C call)."""
    d = ProgramDesc()
    d.dtype, d.flags = dtype, flags
    d.n_inputs = len(in_width)
    iw, im = _i32(in_width), _i32(in_mode)
    d.input_width, d.input_mode = iw, im
    d.n_outputs = len(out_width)
    ow, om = _i32(out_width), _i32(out_mode)
    d.output_width, d.output_mode = ow, om
    d.n_params = n_params
    cs = (C.c_double * max(len(consts), 1))(*consts)
    d.n_consts, d.consts = len(consts), cs
    ins = (Instr * max(len(instr), 1))(*[Instr(*map(int, t)) for t in instr])
    d.n_instr, d.instr = len(instr), ins
    sts = (Store * max(len(stores), 1))(*[Store(*map(int, t)) for t in stores])
    d.n_stores, d.stores = len(stores), sts
    d.vec = vec
    keep.extend([iw, im, ow, om, cs, ins, sts])
    return d


class PipelineProgram:
    """A compiled pipeline (include/numga_b200.h `ngb200_pipeline_*`): several straight-line programs in ONE kernel,
    exchanging values through shared memory.  Built by numga_b200/pipeline.py."""

    def __init__(self, dtype, flags, items_per_tile, globals_, shared, phases, loop, n_params, threads_per_tile=0):
        """globals_: [(width, domain | -1, is_index)]; shared: [(width, domain)];
        phases: [(ir dict, domain, [(kind, slot, index)] inputs, [(kind, slot, index)] outputs)]"""
        keep = []
        d = PipelineDesc()
        d.dtype, d.flags = dtype, flags
        ipt = _i32(items_per_tile)
        d.n_domains, d.items_per_tile = len(items_per_tile), ipt
        gw, gd, gi = _i32([g[0] for g in globals_]), _i32([g[1] for g in globals_]), _i32([int(g[2]) for g in globals_])
        d.n_globals, d.global_width, d.global_domain, d.global_is_index = len(globals_), gw, gd, gi
        sw, sd = _i32([s[0] for s in shared]), _i32([s[1] for s in shared])
        d.n_shared, d.shared_width, d.shared_domain = len(shared), sw, sd
        arr = (PhaseDesc * len(phases))()
        for k, (ir, domain, ins, outs) in enumerate(phases):
            arr[k].program = _program_desc(dtype, flags, ir['in_width'], [VARYING] * len(ir['in_width']), ir['out_width'],
                                           [VARYING] * len(ir['out_width']), ir['n_params'], ir['consts'], ir['instr'],
                                           ir['stores'], 0, keep)
            arr[k].domain = domain
            bi = (Bind * max(len(ins), 1))(*[Bind(*map(int, b)) for b in ins])
            bo = (Bind * max(len(outs), 1))(*[Bind(*map(int, b)) for b in outs])
            arr[k].inputs, arr[k].outputs = bi, bo
            keep.extend([bi, bo])
        d.n_phases, d.phases = len(phases), arr
        d.loop_begin, d.loop_end = loop
        d.n_params, d.threads_per_tile = n_params, threads_per_tile
        h = C.c_void_p()
        check(lib.ngb200_pipeline_create(C.byref(d), C.byref(h)))
        self.handle = h
        self.n_globals, self.n_domains, self.n_params = len(globals_), len(items_per_tile), n_params

    def __del__(self):
        try:
            if self.handle:
                lib.ngb200_pipeline_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def source(self):
        return lib.ngb200_pipeline_source(self.handle).decode()

    @property
    def info(self):
        info = PipelineInfo()
        check(lib.ngb200_pipeline_get_info(self.handle, C.byref(info)))
        return {k: getattr(info, k) for k, _ in PipelineInfo._fields_}

    def launch(self, globals_, domain_items, n_tiles, repeat=1, params=(), stream=0):
        """globals_: [(device_ptr, comp_stride, batch_stride)] per global operand (index arrays: (ptr, 0, 1))."""
        assert len(globals_) == self.n_globals and len(domain_items) == self.n_domains
        p = (C.c_double * max(len(params), 1))(*params)
        n = (C.c_int64 * self.n_domains)(*[int(x) for x in domain_items])
        check(lib.ngb200_pipeline_launch(self.handle, Program._operands(globals_), n, int(n_tiles), int(repeat), p,
                                         C.c_void_p(stream)))


class Program:
    """A compiled straight-line program (one cubin: a contiguous kernel, vectorised where that pays, + a generic
    strided / gather-scatter kernel)."""

    def __init__(self, handle, in_width, in_mode, out_width, out_mode, n_params, dtype):
        self.handle = C.c_void_p(handle)
        self.in_width, self.in_mode = tuple(in_width), tuple(in_mode)
        self.out_width, self.out_mode = tuple(out_width), tuple(out_mode)
        self.n_params = n_params
        self.dtype = dtype

    @classmethod
    def from_ir(cls, dtype, flags, in_width, in_mode, out_width, out_mode, n_params, consts, instr, stores, vec=0):
        """instr: sequence of (op, a, b, c); stores: sequence of (operand, component, value)."""
        keep = []
        d = _program_desc(dtype, flags, in_width, in_mode, out_width, out_mode, n_params, consts, instr, stores, vec, keep)
        h = C.c_void_p()
        check(lib.ngb200_program_create(C.byref(d), C.byref(h)))
        return cls(h.value, in_width, in_mode, out_width, out_mode, n_params, dtype)

    @classmethod
    def from_terms(cls, terms, n_in, n_out, dtype, bcast_mask=0, flags=FAST):
        """terms: int array [n_terms, arity + 2] rows (i_0 .. i_{k-1}, o, coeff) in precompute_sparse order."""
        n_in = list(n_in)
        arity = len(n_in)
        terms = np.ascontiguousarray(np.asarray(terms, dtype=np.int32).reshape(-1, arity + 2))
        h = C.c_void_p()
        check(lib.ngb200_operator_create(
            terms.ctypes.data_as(C.POINTER(C.c_int32)), len(terms), arity, _i32(n_in), n_out, dtype,
            bcast_mask, flags, C.byref(h)))
        modes = [BROADCAST if (bcast_mask >> k) & 1 else VARYING for k in range(arity)]
        return cls(h.value, n_in, modes, [n_out], [VARYING], 0, dtype)

    def __del__(self):
        try:
            if self.handle:
                lib.ngb200_program_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def source(self):
        return lib.ngb200_program_source(self.handle).decode()

    @property
    def info(self):
        info = ProgramInfo()
        check(lib.ngb200_program_get_info(self.handle, C.byref(info)))
        return {k: getattr(info, k) for k, _ in ProgramInfo._fields_}

    @staticmethod
    def _operands(specs):
        arr = (Operand * max(len(specs), 1))()
        for k, (ptr, cs, bs, *rest) in enumerate(specs):
            arr[k] = Operand(ptr, cs, bs, rest[0] if rest else None, rest[1] if len(rest) > 1 else 0,
                             rest[2] if len(rest) > 2 else 0)
        return arr

    def launch(self, inputs, outputs, batch, stream=0, params=()):
        """inputs/outputs: sequences of (device_ptr, comp_stride, batch_stride[, index_ptr[, inner_size, outer_stride]])."""
        assert len(inputs) == len(self.in_width) and len(outputs) == len(self.out_width)
        p = (C.c_double * max(len(params), 1))(*params)
        check(lib.ngb200_program_launch(self.handle, self._operands(inputs), self._operands(outputs), p,
                                        int(batch), C.c_void_p(stream)))

    def run_host(self, inputs, outputs, batch, device=0, chunk_items=0, params=()):
        """Same as launch, but with host pointers; synchronous."""
        p = (C.c_double * max(len(params), 1))(*params)
        check(lib.ngb200_program_run_host(self.handle, self._operands(inputs), self._operands(outputs), p,
                                          int(batch), int(device), int(chunk_items)))


def fill_normal(ptr, n, seed, offset=0, scale=1.0, dtype=F32, stream=0):
    check(lib.ngb200_fill_normal(C.c_void_p(ptr), int(n), int(seed), int(offset), float(scale), dtype, C.c_void_p(stream)))
