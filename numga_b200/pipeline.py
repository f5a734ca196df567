"""Pipelines: several traced functions fused into ONE kernel that exchanges values on chip.

The eager / `jit` paths turn one expression into one kernel whose items are independent.  A pipeline covers the next
case up: a CHAIN of data-parallel steps over a few item domains with local coupling between them -- the rigid-body
sub-step of numga/examples/physics/core.py:80-88 (`pre_integrate` per body -> `Constraint.apply` per constraint of each
colour set, reading and writing the two bodies it joins -> `post_integrate` per body -> `Constraint.v_apply` ...).
Items are grouped into tiles (e.g. 10 chains of 12 bodies with their 60 + 50 constraints); one CTA owns a tile, keeps
the evolving state in shared memory, runs the phases with CTA barriers in between and touches HBM only for the
initial load, the constants and the final store.  The phases are the SAME python bodies the other paths trace.

    pipe = Pipeline(ctx, {'body': 120, 'red': 60, 'black': 50})
    motor   = pipe.array('motor', S.even_grade(), 'body')          # global (HBM) operand
    i0      = pipe.index('red_i0', 'red')                          # int64 index array over the 'red' domain
    s_motor = pipe.buffer(S.even_grade(), 'body')                  # on-chip buffer
    pipe.phase('body', load_fn, [motor], [s_motor])
    pipe.phase('red', apply_fn, [s_motor[i0], s_motor[i1], ..., pipe.param(0)], [s_motor[i0], s_motor[i1]])
    compiled = pipe.build(loop=(1, 7))
    compiled.launch({'motor': mv, 'red_i0': idx, ...}, {'body': n_bodies, ...}, repeat=10, params=[dt])

C ABI: `ngb200_pipeline_*` (include/numga_b200.h).
"""
import numpy as np
import torch

from numga_b200 import runtime as rt


def operator_values(op):
    """Inside a traced phase: the kernel entries of a bound operator argument as a multivector over its kernel space,
    so that a phase can RETURN them (e.g. to park an inverse-inertia table in shared memory)."""
    from numga_b200.backend import kernel_space
    from numga_b200.multivector import TracedMultiVector
    return TracedMultiVector(op.context, kernel_space(op.kshape), op._flat)


def compact_operator(op, mask=None):
    """Inside a traced phase: the structurally non-zero kernel entries of a bound operator (C order over `mask`) as a
    multivector over a compact kernel space -- what gets parked in shared memory."""
    from numga_b200.backend import kernel_space
    from numga_b200.multivector import TracedMultiVector
    from numga_b200.trace import SymArray
    mask = op.mask if mask is None else mask
    pos = np.flatnonzero(np.asarray(mask).reshape(-1))
    regs = op._flat.regs
    return TracedMultiVector(op.context, kernel_space((len(pos),)), SymArray(op._flat.trace, [regs[p] for p in pos]))


def expand_operator(context, axes, mask, compact):
    """Inverse of `compact_operator`: a bound operator over `axes` whose masked entries are the components of the traced
    multivector `compact` and whose other entries are the constant 0."""
    from numga_b200.backend import B200DenseOperator
    from numga_b200.trace import SymArray
    trace = compact.values.trace
    zero = trace.const(0.0)
    flat = [zero] * int(np.prod(np.asarray(mask).shape))
    for j, p in enumerate(np.flatnonzero(np.asarray(mask).reshape(-1))):
        flat[p] = compact.values.regs[j]
    return B200DenseOperator(context, axes, SymArray(trace, flat), np.asarray(mask))


class Ref:
    """A pipeline operand: a global array / index array / shared buffer, optionally addressed through an index."""

    def __init__(self, pipe, kind, slot, space, domain, index=None, name=None):
        self.pipe, self.kind, self.slot, self.space, self.domain, self.index, self.name = pipe, kind, slot, space, domain, index, name

    def __getitem__(self, index):
        assert isinstance(index, Ref) and index.kind == 'index', 'pipeline operands are indexed by pipe.index(...) arrays'
        assert self.kind in ('global', 'shared') and self.index is None
        return Ref(self.pipe, self.kind, self.slot, self.space, self.domain, index, self.name)

    def bind(self):
        return (rt.BIND_SHARED if self.kind == 'shared' else rt.BIND_GLOBAL, self.slot, -1 if self.index is None else self.index.slot)


class _Param:
    def __init__(self, k):
        self.k = k


class Pipeline:
    def __init__(self, context, items_per_tile, threads_per_tile=0):
        self.context = context
        self.domains = list(items_per_tile)
        self.items_per_tile = [int(items_per_tile[d]) for d in self.domains]
        self.threads_per_tile = threads_per_tile
        self.globals, self.shared, self.phases = [], [], []
        self.n_params = 0

    def _domain(self, name):
        return -1 if name is None else self.domains.index(name)

    def array(self, name, space, domain):
        """Global (HBM) array of multivectors over `space` (a SubSpace, or the kernel key of a bound operator), one per
        item of `domain`; domain None = ONE multivector shared by every item (read-only)."""
        self.globals.append((name, len(space), self._domain(domain), False))
        return Ref(self, 'global', len(self.globals) - 1, space, self._domain(domain), name=name)

    def index(self, name, domain):
        self.globals.append((name, 1, self._domain(domain), True))
        return Ref(self, 'index', len(self.globals) - 1, None, self._domain(domain), name=name)

    def buffer(self, space, domain):
        self.shared.append((len(space), self._domain(domain)))
        return Ref(self, 'shared', len(self.shared) - 1, space, self._domain(domain))

    def param(self, k):
        self.n_params = max(self.n_params, k + 1)
        return _Param(k)

    def phase(self, domain, fn, inputs, outputs):
        """fn(*multivectors_and_scalars) -> multivector(s), traced once; `inputs` are Refs or pipe.param(k) in fn's
        argument order (the k-th DISTINCT param object becomes the program's k-th scalar), `outputs` Refs per result."""
        key = []
        for a in inputs:
            if isinstance(a, _Param):
                key.append(('param',))
            else:
                key.append(('mv', a.space, rt.BROADCAST if (a.kind == 'global' and a.domain < 0) else rt.VARYING))
        ir, subspaces, single, roles = self.context.engine.trace(fn, tuple(key))
        params = [a.k for a in inputs if isinstance(a, _Param)]
        assert params == list(range(len(params))), 'scalar params must be passed in pipeline order (param(0), param(1), ...)'
        assert len(subspaces) == len(outputs), f'{len(outputs)} destinations for {len(subspaces)} results'
        for s, o in zip(subspaces, outputs):
            assert len(s) == len(o.space), f'result over {s!r} does not fit destination over {o.space!r}'
        refs = [a for a in inputs if not isinstance(a, _Param)]
        self.phases.append((ir, self._domain(domain), refs, list(outputs)))
        return len(self.phases) - 1

    def build(self, loop=(0, 0)):
        return CompiledPipeline(self, loop)


class CompiledPipeline:
    def __init__(self, pipe, loop):
        ctx = pipe.context
        self.context, self.pipe, self.loop = ctx, pipe, tuple(loop)
        self.dtype = rt.F32 if ctx.dtype == torch.float32 else rt.F64
        self.phases = [(ir, d, [r.bind() for r in ins], [r.bind() for r in outs]) for ir, d, ins, outs in pipe.phases]
        self.program = rt.PipelineProgram(
            self.dtype, rt.FAITHFUL if ctx.faithful else rt.FAST, pipe.items_per_tile,
            [(w, d, i) for _, w, d, i in pipe.globals], pipe.shared, self.phases, self.loop, pipe.n_params,
            pipe.threads_per_tile)

    def operands(self, arrays):
        """(keep-alive list, [(ptr, comp_stride, batch_stride)]) for the global operands, by name."""
        from numga_b200.backend import B200DenseOperator, flat_operand, _prod
        ops, keep = [], []
        for name, width, domain, is_index in self.pipe.globals:
            a = arrays[name]
            if is_index:
                assert a.dtype == torch.int64 and a.is_contiguous(), f'{name}: index arrays are contiguous int64 tensors'
                keep.append(a)
                ops.append((a.data_ptr(), 0, 1))
                continue
            v = a.flat if isinstance(a, B200DenseOperator) else a.values
            assert v.shape[-1] == width, f'{name}: {tuple(v.shape)} does not hold {width} components'
            t, ptr, cs, bs = flat_operand(v, _prod(v.shape[:-1]))
            keep.append(t)
            ops.append((ptr, cs, 1 if domain >= 0 else 0))
            assert domain < 0 or bs == 1 or _prod(v.shape[:-1]) == 1, f'{name}: not structure-of-arrays'
        return keep, ops

    def _check_aliasing(self, tensors):
        """Operands that no phase writes are read through the non-coherent path (`ld.global.nc`) by the kernel, which
        is only valid if nothing else in the launch writes that memory: a written operand may not overlap a read-only one."""
        written = {b[1] for _, _, _, outs in self.phases for b in outs if b[0] == rt.BIND_GLOBAL}
        spans = []
        for slot, t in enumerate(tensors):
            st = t.untyped_storage()
            spans.append((st.data_ptr(), st.data_ptr() + st.nbytes(), slot))
        for lo, hi, slot in spans:
            if slot not in written:
                continue
            for lo2, hi2, other in spans:
                if other != slot and other not in written and lo < hi2 and lo2 < hi and hi > lo and hi2 > lo2:
                    raise ValueError(f'pipeline output {self.pipe.globals[slot][0]!r} shares memory with the read-only operand '
                                     f'{self.pipe.globals[other][0]!r}; pass a separate destination')

    def validate(self, arrays, items):
        """Bounds check of every index operand (the kernel itself does none): rows addressed in a shared buffer must
        fall inside the tile of the item that addresses them, rows addressed in a global array inside that array.  Done
        once per index tensor (identity + version) from `launch`; costs a few small torch reductions."""
        pipe = self.pipe
        counts = [int(items[d]) for d in pipe.domains]
        seen = self.__dict__.setdefault('_validated', set())
        for ir, d, ins, outs in self.phases:
            for kind, slot, index in list(ins) + list(outs):
                if index < 0:
                    continue
                name = pipe.globals[index][0]
                idx = arrays[name]
                target = pipe.shared[slot][1] if kind == rt.BIND_SHARED else pipe.globals[slot][2]
                key = (name, kind, target, idx.data_ptr(), idx._version, counts[d], counts[target])
                if key in seen:
                    continue
                n = counts[d]
                assert idx.numel() >= n, f'{name}: {idx.numel()} indices for {n} items'
                rows = idx.reshape(-1)[:n]
                ok = (rows >= 0) & (rows < counts[target])
                if kind == rt.BIND_SHARED:
                    ipt_src, ipt_dst = pipe.items_per_tile[d], pipe.items_per_tile[target]
                    tile = torch.div(torch.arange(n, device=rows.device), ipt_src, rounding_mode='floor')
                    local = rows - tile * ipt_dst
                    ok &= (local >= 0) & (local < ipt_dst)
                if n and not bool(ok.all()):
                    raise ValueError(f'pipeline index operand {name!r} addresses rows outside its tile / array '
                                     f'(first offender: item {int((~ok).nonzero()[0])})')
                seen.add(key)

    def launch(self, arrays, items, repeat=1, params=()):
        """arrays: {name: multivector | bound operator | int64 tensor}; items: {domain: total item count}."""
        ctx, pipe = self.context, self.pipe
        self.validate(arrays, items)
        counts = [int(items[d]) for d in pipe.domains]
        n_tiles = max(-(-c // ipt) for c, ipt in zip(counts, pipe.items_per_tile))
        keep, ops = self.operands(arrays)
        self._check_aliasing(keep)
        if n_tiles > 0 and ctx.can_launch():
            self.program.launch(ops, counts, n_tiles, repeat, [float(p) for p in params], ctx.stream_handle())
        return keep
