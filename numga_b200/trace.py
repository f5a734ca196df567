"""Symbolic tracing of multivector code into the straight-line IR of include/numga_b200.h.

`SymArray` stands in for the `values` array of a multivector while a function is traced: it has
one SSA value per component and supports the arithmetic the multivector layer performs on
`values` (`+ - * / **`, unary minus, `[..., i]`).  Batch axes are implicit: every traced
operation is elementwise over the batch, which is exactly the structure of the hot path
(numga/backend/numpy/operator.py:117-132 and the elementwise glue of
numga/multivector/multivector.py:247-261,307-313,398-411).

The `Trace` hash-conses instructions (common sub-expressions are shared) and applies only
EXACT rewrites (x*1, x+0, -(-x), a-(-b) = a+b, constant folding in the program dtype), so the
"faithful" programs stay bit-identical to evaluating the same chain op by op.
"""
import numpy as np

from numga_b200 import runtime as rt


class Trace:
    def __init__(self, np_dtype, faithful=False):
        self.np_dtype = np.dtype(np_dtype)
        self.faithful = faithful
        self.instr = []          # (op, a, b, c)
        self.consts = []         # python floats, already rounded to dtype
        self.uniform = []        # per value: depends only on broadcast inputs / consts / params
        self.memo = {}
        self.input_width, self.input_mode = [], []
        self.n_params = 0

    # ---- leaves ---------------------------------------------------------------------------------
    def _push(self, key, uniform):
        if key in self.memo:
            return self.memo[key]
        self.instr.append(key)
        self.uniform.append(uniform)
        self.memo[key] = len(self.instr) - 1
        return self.memo[key]

    def add_input(self, width, mode):
        k = len(self.input_width)
        self.input_width.append(int(width))
        self.input_mode.append(int(mode))
        return [self._push((rt.OP_INPUT, k, c, 0), mode == rt.BROADCAST) for c in range(width)]

    def add_param(self):
        k = self.n_params
        self.n_params += 1
        return self._push((rt.OP_PARAM, k, 0, 0), True)

    def const(self, value):
        v = float(self.np_dtype.type(value))
        key = ('const', v if v == v else 'nan', np.signbit(v))
        if key in self.memo:
            return self.memo[key]
        self.consts.append(v)
        vid = self._push((rt.OP_CONST, len(self.consts) - 1, 0, 0), True)
        self.memo[key] = vid
        return vid

    def const_value(self, vid):
        op, a, _, _ = self.instr[vid]
        return self.consts[a] if op == rt.OP_CONST else None

    def is_const(self, vid, value):
        c = self.const_value(vid)
        return c is not None and c == value and not (value == 0 and np.signbit(c))

    def _neg_of(self, vid):
        op, a, _, _ = self.instr[vid]
        return a if op == rt.OP_NEG else None

    def _fold(self, fn, *vids):
        vals = [self.const_value(v) for v in vids]
        if any(v is None for v in vals):
            return None
        t = self.np_dtype.type
        with np.errstate(all='ignore'):
            return self.const(fn(*[t(v) for v in vals]))

    def _emit(self, op, a=0, b=0, c=0):
        n = rt.N_OPERANDS.get(op, 2)
        u = all(self.uniform[x] for x in (a, b, c)[:n])
        return self._push((op, a, b, c), u)

    # ---- arithmetic with exact simplifications ------------------------------------------------------
    def neg(self, a):
        f = self._fold(lambda x: -x, a)
        if f is not None:
            return f
        inner = self._neg_of(a)
        return inner if inner is not None else self._emit(rt.OP_NEG, a)

    def _is_zero(self, vid):
        """+0.0 or -0.0 (x + (-0.0) == x exactly; x + (+0.0) differs from x only in the sign of a zero)"""
        c = self.const_value(vid)
        return c is not None and c == 0

    def add(self, a, b):
        f = self._fold(lambda x, y: x + y, a, b)
        if f is not None:
            return f
        if self._is_zero(a):
            return b
        if self._is_zero(b):
            return a
        nb = self._neg_of(b)
        if nb is not None:
            return self.sub(a, nb)
        na = self._neg_of(a)
        if na is not None:
            return self.sub(b, na)
        return self._emit(rt.OP_ADD, min(a, b), max(a, b))

    def sub(self, a, b):
        f = self._fold(lambda x, y: x - y, a, b)
        if f is not None:
            return f
        if self._is_zero(b):
            return a
        if self._is_zero(a):
            return self.neg(b)
        nb = self._neg_of(b)
        if nb is not None:
            return self.add(a, nb)
        return self._emit(rt.OP_SUB, a, b)

    def mul(self, a, b):
        f = self._fold(lambda x, y: x * y, a, b)
        if f is not None:
            return f
        for x, y in ((a, b), (b, a)):
            if self.is_const(x, 1.0):
                return y
            if self.is_const(x, -1.0):
                return self.neg(y)
            if self.is_const(x, 0.0):
                return x
        for x, y in ((a, b), (b, a)):
            cv = self.const_value(x)
            if cv is not None and cv < 0:
                # (-c) * y = -(c * y) exactly: 2y and -2y then share one multiply, and the sign folds into the
                # operand of the consuming add / FMA for free
                return self.neg(self.mul(self.const(-cv), y))
        na, nb = self._neg_of(a), self._neg_of(b)
        if na is not None and nb is not None:
            return self.mul(na, nb)
        if na is not None:
            return self.neg(self.mul(na, b))
        if nb is not None:
            return self.neg(self.mul(a, nb))
        return self._emit(rt.OP_MUL, min(a, b), max(a, b))

    def div(self, a, b):
        f = self._fold(lambda x, y: x / y, a, b)
        if f is not None:
            return f
        if self.is_const(b, 1.0):
            return a
        if not self.faithful:
            # FAST programs: a / b = a * (1 / b).  The reciprocal is hash-consed, so the components of a multivector
            # divided by one scalar share ONE reciprocal, and the multiplies fuse into FMAs downstream.
            return self.mul(a, self.rcp(b))
        if self.is_const(a, 1.0):
            return self._emit(rt.OP_RCP, b)
        return self._emit(rt.OP_DIV, a, b)

    def rcp(self, b):
        """1 / b with the exact-enough rewrites of FAST mode: 1/(c*y) = (1/c) * (1/y), 1/sqrt(x) = rsqrt(x)."""
        f = self._fold(lambda x: 1 / x, b)
        if f is not None:
            return f
        op, x, y, _ = self.instr[b]
        if not self.faithful:
            if op == rt.OP_SQRT:
                return self._emit(rt.OP_RSQRT, x)
            if op == rt.OP_MUL:
                for c, other in ((x, y), (y, x)):
                    cv = self.const_value(c)
                    if cv is not None and cv != 0 and np.isfinite(cv):
                        return self.mul(self.const(1.0 / cv), self.rcp(other))
            if op == rt.OP_NEG:
                return self.neg(self.rcp(x))
        return self._emit(rt.OP_RCP, b)

    def unary(self, op, a, fn=None):
        if fn is not None:
            f = self._fold(fn, a)
            if f is not None:
                return f
        return self._emit(op, a)

    def sqrt(self, a):
        return self.unary(rt.OP_SQRT, a, np.sqrt)

    def rsqrt(self, a):
        return self.unary(rt.OP_RSQRT, a, lambda x: 1 / np.sqrt(x))

    def abs(self, a):
        return self.unary(rt.OP_ABS, a, np.abs)

    def nan_to_num(self, a, nan):
        return self._emit(rt.OP_NAN_TO_NUM, a, self.const(nan))

    def atan2(self, y, x):
        f = self._fold(np.arctan2, y, x)
        return f if f is not None else self._emit(rt.OP_ATAN2, y, x)

    def partner(self, a):
        """The value `a` as computed by the thread of the paired item (i ^ 1): pipeline phases only (NGB200_OP_PARTNER)."""
        return a if self.uniform[a] else self._emit(rt.OP_PARTNER, a)

    def select_gt0(self, cond, a, b):
        """cond > 0 ? a : b"""
        if a == b:
            return a
        c = self.const_value(cond)
        if c is not None:
            return a if c > 0 else b
        return self._emit(rt.OP_SELECT_GT0, cond, a, b)

    def power(self, a, p):
        if p == 0.5:
            return self.sqrt(a)
        if p == -0.5:
            return self.rsqrt(a)
        if p == 1:
            return a
        if p == -1:
            return self.rcp(a)
        if p == 2:
            return self.mul(a, a)
        raise NotImplementedError(f'power {p} is not supported in traced code')

    # ---- output ----------------------------------------------------------------------------------------
    def finish(self, outputs):
        """outputs: list of lists of value ids (one list per output multivector). Returns IR pieces
        with dead code removed and values renumbered."""
        live = set()
        stack = [v for out in outputs for v in out]
        while stack:
            v = stack.pop()
            if v in live:
                continue
            live.add(v)
            op, a, b, c = self.instr[v]
            stack.extend((a, b, c)[:rt.N_OPERANDS.get(op, 2)])
        order = sorted(live)
        renum = {v: k for k, v in enumerate(order)}
        used_consts, instr = {}, []
        for v in order:
            op, a, b, c = self.instr[v]
            n = rt.N_OPERANDS.get(op, 2)
            if op == rt.OP_CONST:
                a = used_consts.setdefault(a, len(used_consts))
            elif n:
                a, b, c = [renum[x] if k < n else 0 for k, x in enumerate((a, b, c))]
            instr.append((op, a, b, c))
        consts = [None] * len(used_consts)
        for old, new in used_consts.items():
            consts[new] = self.consts[old]
        stores = [(k, c, renum[v]) for k, out in enumerate(outputs) for c, v in enumerate(out)]
        return dict(instr=instr, consts=consts, stores=stores, out_width=[len(o) for o in outputs],
                    in_width=list(self.input_width), in_mode=list(self.input_mode), n_params=self.n_params)


class SymArray:
    """Symbolic stand-in for a multivector's `values` array ([..., n], batch axes implicit)."""

    def __init__(self, trace: Trace, regs):
        self.trace = trace
        self.regs = list(regs)

    @property
    def shape(self):
        return (Ellipsis, len(self.regs))

    def __len__(self):
        return len(self.regs)

    def copy(self):
        return SymArray(self.trace, self.regs)

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            if len(idx) == 2 and idx[0] is Ellipsis:
                idx = idx[1]
            else:
                raise IndexError('traced values support [..., i] / [..., a:b] only')
        if idx is None:
            return self
        if isinstance(idx, slice):
            return SymArray(self.trace, self.regs[idx])
        if isinstance(idx, (list, np.ndarray)):
            return SymArray(self.trace, [self.regs[int(i)] for i in idx])
        return SymArray(self.trace, [self.regs[int(idx)]])

    def _other(self, other):
        if isinstance(other, SymArray):
            assert other.trace is self.trace
            regs = other.regs
        elif isinstance(other, (int, float, np.floating, np.integer)):
            regs = [self.trace.const(other)]
        else:
            arr = np.asarray(other)
            if arr.dtype == object:
                return None
            regs = [self.trace.const(v) for v in arr.reshape(-1)]
        if len(self.regs) == 0 and len(regs) <= 1 or len(regs) == 0 and len(self.regs) <= 1:
            return [], []                       # an empty multivector (no components) stays empty
        n = max(len(regs), len(self.regs))
        mine = self.regs * n if len(self.regs) == 1 else self.regs
        regs = regs * n if len(regs) == 1 else regs
        assert len(mine) == len(regs) == n, 'component counts do not broadcast'
        return mine, regs

    def _binary(self, other, fn, swap=False):
        pair = self._other(other)
        if pair is None:
            return NotImplemented
        a, b = pair
        if swap:
            a, b = b, a
        return SymArray(self.trace, [fn(x, y) for x, y in zip(a, b)])

    def __add__(self, o): return self._binary(o, self.trace.add)
    def __radd__(self, o): return self._binary(o, self.trace.add, swap=True)
    def __sub__(self, o): return self._binary(o, self.trace.sub)
    def __rsub__(self, o): return self._binary(o, self.trace.sub, swap=True)
    def __mul__(self, o): return self._binary(o, self.trace.mul)
    def __rmul__(self, o): return self._binary(o, self.trace.mul, swap=True)
    def __truediv__(self, o): return self._binary(o, self.trace.div)
    def __rtruediv__(self, o): return self._binary(o, self.trace.div, swap=True)
    def __neg__(self): return SymArray(self.trace, [self.trace.neg(r) for r in self.regs])
    def __pos__(self): return self
    def __pow__(self, p): return SymArray(self.trace, [self.trace.power(r, p) for r in self.regs])
    def __abs__(self): return SymArray(self.trace, [self.trace.abs(r) for r in self.regs])

    def map(self, fn):
        return SymArray(self.trace, [fn(r) for r in self.regs])


# ---- lowering of a symbolic operator applied to traced operands ---------------------------------------

def lower_operator(trace: Trace, operator, inputs):
    """Append `operator(*inputs)` to the trace; inputs are lists of value ids per argument.
    Returns the value ids of the output components (implicit zeros for rows without terms)."""
    zero = None
    out = []
    rows = dict(operator.precompute_sparse)
    for o in range(len(operator.output)):
        terms = rows.get(o)
        if not terms:
            zero = trace.const(0.0) if zero is None else zero
            out.append(zero)
        elif trace.faithful:
            out.append(_row_faithful(trace, terms, inputs))
        else:
            out.append(_row_fast(trace, [(c, [inputs[k][i] for k, i in enumerate(idx)]) for idx, c in terms]))
    return out


def _row_faithful(trace, terms, inputs):
    """Terms in precompute_sparse order; each product rounded factor by factor, then a sequential sum:
    the order in which numpy's einsum accumulates (hard part 1)."""
    acc = None
    for idx, c in terms:
        prod = None
        for k in reversed(range(len(idx))):     # right-associated: a0 * (a1 * a2), einsum's contraction order
            prod = inputs[k][idx[k]] if prod is None else trace.mul(inputs[k][idx[k]], prod)
        if abs(c) != 1:
            prod = trace.mul(prod, trace.const(abs(c)))
        if acc is None:
            acc = prod if c > 0 else trace.neg(prod)
        else:
            acc = trace.add(acc, prod) if c > 0 else trace.sub(acc, prod)
    return acc


def _row_fast(trace, terms):
    """sum_t c_t * prod(factors_t), arranged for few multiplies:
      * factors that are uniform over the batch are multiplied first and their partial sums are
        shared, so a broadcast operand is pre-contracted once per thread (sandwich -> 4x4 matrix);
      * products are built in a canonical order so twin terms of a repeated argument
        (R_i R_k and R_k R_i) become one value with a summed integer coefficient;
      * the sum is factored over the last remaining varying factor (Horner-like)."""
    # merge identical products
    merged = {}
    for c, factors in terms:
        key = tuple(sorted(factors))
        merged[key] = merged.get(key, 0) + c
    groups = {}   # varying factor tuple -> list of (coeff, uniform factor tuple)
    for key, c in merged.items():
        if c == 0:
            continue
        uni = tuple(f for f in key if trace.uniform[f])
        var = tuple(f for f in key if not trace.uniform[f])
        groups.setdefault(var, []).append((c, uni))

    def signed_sum(items):
        """items: list of (int coeff, value id or None for 1) -> value id"""
        acc = None
        for c, v in sorted(items, key=lambda t: (t[1] is None, t[1] if t[1] is not None else 0)):
            if v is None:
                term = trace.const(c)
                sign = 1
            else:
                term = v if abs(c) == 1 else trace.mul(v, trace.const(abs(c)))
                sign = 1 if c > 0 else -1
            if acc is None:
                acc = term if sign > 0 else trace.neg(term)
            else:
                acc = trace.add(acc, term) if sign > 0 else trace.sub(acc, term)
        return acc

    def product(factors):
        acc = None
        for f in factors:
            acc = f if acc is None else trace.mul(acc, f)
        return acc

    # one coefficient (uniform) per varying product
    partial = []
    for var, items in groups.items():
        coef = signed_sum([(c, product(uni) if uni else None) for c, uni in items])
        partial.append((var, coef))
    # factor over the least-shared varying factor: pick the factor position whose removal leaves
    # the most repeated remainders -> here: group by the last factor (highest value id)
    by_pivot = {}
    for var, coef in partial:
        if not var:
            by_pivot.setdefault(None, []).append(((), coef))
        else:
            by_pivot.setdefault(var[-1], []).append((var[:-1], coef))
    acc = None
    for pivot in sorted(by_pivot, key=lambda p: -1 if p is None else p):
        inner = None
        for rest, coef in by_pivot[pivot]:
            t = coef if not rest else trace.mul(product(rest), coef)
            inner = t if inner is None else trace.add(inner, t)
        term = inner if pivot is None else trace.mul(inner, pivot)
        acc = term if acc is None else trace.add(acc, term)
    return trace.const(0.0) if acc is None else acc        # every term cancelled (e.g. x ^ x on one traced value)


def lower_partial(trace: Trace, operator, bound, free, kshape):
    """Kernel of `operator` with inputs `bound` ({axis: value ids}) substituted: for every remaining
    (free input indices..., output) entry, sum coeff * prod(bound factors).  Returns the flattened kernel
    (C order over kshape); entries without terms are the constant 0."""
    arity = operator.arity
    groups = {}
    for idx, c in zip(operator.index.tolist(), operator.coeff.tolist()):
        key = tuple(idx[k] for k in free) + (idx[arity],)
        groups.setdefault(key, []).append((c, [bound[k][idx[k]] for k in sorted(bound)]))
    zero = trace.const(0.0)
    size = 1
    for n in kshape:
        size *= n
    flat = [zero] * size
    for key, terms in groups.items():
        pos = 0
        for k, n in zip(key, kshape):
            pos = pos * n + k
        if trace.faithful:
            acc = None
            for c, factors in terms:
                prod = None
                for f in reversed(factors):
                    prod = f if prod is None else trace.mul(f, prod)
                if abs(c) != 1:
                    prod = trace.mul(prod, trace.const(abs(c)))
                if acc is None:
                    acc = prod if c > 0 else trace.neg(prod)
                else:
                    acc = trace.add(acc, prod) if c > 0 else trace.sub(acc, prod)
            flat[pos] = acc
        else:
            flat[pos] = _row_fast(trace, terms)
    return flat


def lower_dense(trace: Trace, kernel, kshape, mask, inputs):
    """out[o] = sum over input index tuples of kernel[j.., o] * prod(inputs[k][j_k]); only entries whose
    `mask` is set are visited.  `kernel` is the flattened (C order) list of value ids."""
    import itertools
    n_out = kshape[-1]
    zero = None
    out = []
    for o in range(n_out):
        acc = None
        for js in itertools.product(*[range(n) for n in kshape[:-1]]):
            if not mask[js + (o,)]:
                continue
            pos = 0
            for k, n in zip(js + (o,), kshape):
                pos = pos * n + k
            prod = kernel[pos]
            for k, j in enumerate(js):
                prod = trace.mul(prod, inputs[k][j])
            acc = prod if acc is None else trace.add(acc, prod)
        if acc is None:
            zero = trace.const(0.0) if zero is None else zero
            acc = zero
        out.append(acc)
    return out


def lower_inverse(trace: Trace, kernel, n, mask=None):
    """Inverse of an n x n matrix given as the flattened (row-major) list of value ids: Gauss-Jordan elimination on
    [A | I] with PARTIAL PIVOTING, fully unrolled into the straight-line IR (row swaps are component-wise selects on
    |a_ik| - |a_kk| > 0), so a batch of small matrices is inverted by one generated kernel, one matrix per thread,
    everything in registers.  The counterpart of `np.linalg.inv` on operator kernels (numga/operator/operator.py:78-90,
    backend/numpy/operator.py `inverse`); same pivoting rule as LAPACK's getrf, results agree to rounding.
    Entries whose `mask` is False are structural zeros.  Returns the flattened inverse (row-major)."""
    zero, one = trace.const(0.0), trace.const(1.0)
    a = [[kernel[i * n + j] if mask is None or mask[i][j] else zero for j in range(n)] for i in range(n)]
    b = [[one if i == j else zero for j in range(n)] for i in range(n)]
    for k in range(n):
        for i in range(k + 1, n):                       # bubble the largest pivot candidate into row k
            cond = trace.sub(trace.abs(a[i][k]), trace.abs(a[k][k]))
            if trace.const_value(cond) is not None and not trace.const_value(cond) > 0:
                continue
            for rows in (a, b):
                lo = k if rows is a else 0
                for j in range(lo, n):
                    rk, ri = rows[k][j], rows[i][j]
                    rows[k][j] = trace.select_gt0(cond, ri, rk)
                    rows[i][j] = trace.select_gt0(cond, rk, ri)
        p = trace.rcp(a[k][k])
        for j in range(k + 1, n):
            a[k][j] = trace.mul(a[k][j], p)
        for j in range(n):
            b[k][j] = trace.mul(b[k][j], p)
        for i in range(n):
            if i == k:
                continue
            f = a[i][k]
            if trace.is_const(f, 0.0):
                continue
            for j in range(k + 1, n):
                a[i][j] = trace.sub(a[i][j], trace.mul(f, a[k][j]))
            for j in range(n):
                b[i][j] = trace.sub(b[i][j], trace.mul(f, b[k][j]))
    return [b[i][j] for i in range(n) for j in range(n)]


# ---- reverse-mode differentiation of a program -----------------------------------------------------------------
def build_vjp(ir, np_dtype, faithful=False, wanted=None):
    """Vector-Jacobian product of a straight-line program as ANOTHER straight-line program (one kernel):

        inputs   the original inputs (same addressing modes), then one cotangent operand per original output
        outputs  one operand per original input (or per input listed in `wanted`):
                 sum_o cotangent[o] * d out[o] / d input   (per batch item)

    The forward values are recomputed inside the backward kernel (registers are cheaper than HBM: nothing is saved
    between the passes), then the adjoints are accumulated in reverse program order.  A multilinear operator's
    backward is thus the same integer table contracted on another axis (numga/operator/operator.py:59-65
    `swapaxes`, numga/backend/jax/operator.py:128-139), and the non-linear chains (exp, log, normalized ...) get
    their exact derivative.  Scalar launch parameters are treated as constants."""
    t = Trace(np_dtype, faithful=faithful)
    n_in = len(ir['in_width'])
    in_regs = [t.add_input(w, m if m != rt.INDEXED else rt.VARYING) for w, m in zip(ir['in_width'], ir['in_mode'])]
    params = [t.add_param() for _ in range(ir['n_params'])]
    cot_regs = [t.add_input(w, rt.VARYING) for w in ir['out_width']]
    # forward replay
    new = []
    for op, a, b, c in ir['instr']:
        if op == rt.OP_INPUT:
            v = in_regs[a][b]
        elif op == rt.OP_CONST:
            v = t.const(ir['consts'][a])
        elif op == rt.OP_PARAM:
            v = params[a]
        elif op == rt.OP_ADD: v = t.add(new[a], new[b])
        elif op == rt.OP_SUB: v = t.sub(new[a], new[b])
        elif op == rt.OP_MUL: v = t.mul(new[a], new[b])
        elif op == rt.OP_DIV: v = t.div(new[a], new[b])
        elif op == rt.OP_NEG: v = t.neg(new[a])
        elif op == rt.OP_SQRT: v = t.sqrt(new[a])
        elif op == rt.OP_RSQRT: v = t.rsqrt(new[a])
        elif op == rt.OP_RCP: v = t.rcp(new[a])
        elif op == rt.OP_ABS: v = t.abs(new[a])
        elif op == rt.OP_NAN_TO_NUM: v = t._emit(op, new[a], new[b])
        elif op in (rt.OP_FMA, rt.OP_SELECT_GT0): v = t._emit(op, new[a], new[b], new[c])
        elif op in (rt.OP_MIN, rt.OP_MAX, rt.OP_ATAN2): v = t._emit(op, new[a], new[b])
        else: v = t._emit(op, new[a])          # exp, log, sin, cos
        new.append(v)
    # reverse sweep
    adj = {}

    def acc(v, g):
        if t.const_value(v) is not None:
            return
        adj[v] = g if v not in adj else t.add(adj[v], g)
    for k, c, v in ir['stores']:
        acc(new[v], cot_regs[k][c])
    zero, half = t.const(0.0), t.const(0.5)
    # Adjoints are keyed by NEW ids (hash-consing and the simplifier can merge several old values into one).  Operands
    # always have smaller ids than their consumers, so descending id order is a reverse topological order: every
    # consumer of v has contributed to adj[v] before v is expanded, whatever the old instruction order was.
    n_forward = len(t.instr)
    for v in range(n_forward - 1, -1, -1):
        if v not in adj:
            continue
        op, a, b, c = t.instr[v]
        g = adj[v]
        if op in (rt.OP_INPUT, rt.OP_CONST, rt.OP_PARAM):
            continue
        if op == rt.OP_ADD:
            acc(a, g); acc(b, g)
        elif op == rt.OP_SUB:
            acc(a, g); acc(b, t.neg(g))
        elif op == rt.OP_MUL:
            acc(a, t.mul(g, b)); acc(b, t.mul(g, a))
        elif op == rt.OP_DIV:
            q = t.div(g, b)
            acc(a, q); acc(b, t.neg(t.mul(q, v)))
        elif op == rt.OP_NEG:
            acc(a, t.neg(g))
        elif op == rt.OP_SQRT:
            acc(a, t.mul(t.mul(g, half), t.rcp(v)))
        elif op == rt.OP_RSQRT:
            acc(a, t.neg(t.mul(t.mul(g, half), t.mul(v, t.mul(v, v)))))
        elif op == rt.OP_RCP:
            acc(a, t.neg(t.mul(g, t.mul(v, v))))
        elif op == rt.OP_ABS:
            # d|x|/dx = sign(x), and 0 at x == 0 (torch's convention): g if x > 0, else (-g if -x > 0 else 0)
            acc(a, t._emit(rt.OP_SELECT_GT0, a, g, t._emit(rt.OP_SELECT_GT0, t.neg(a), t.neg(g), zero)))
        elif op == rt.OP_NAN_TO_NUM:
            acc(a, g)
        elif op == rt.OP_FMA:
            acc(a, t.mul(g, b)); acc(b, t.mul(g, a)); acc(c, g)
        elif op == rt.OP_EXP:
            acc(a, t.mul(g, v))
        elif op == rt.OP_LOG:
            acc(a, t.div(g, a))
        elif op == rt.OP_SIN:
            acc(a, t.mul(g, t._emit(rt.OP_COS, a)))
        elif op == rt.OP_COS:
            acc(a, t.neg(t.mul(g, t._emit(rt.OP_SIN, a))))
        elif op in (rt.OP_MIN, rt.OP_MAX):
            d = t.sub(b, a) if op == rt.OP_MIN else t.sub(a, b)       # > 0: a is selected
            acc(a, t._emit(rt.OP_SELECT_GT0, d, g, zero)); acc(b, t._emit(rt.OP_SELECT_GT0, d, zero, g))
        elif op == rt.OP_SELECT_GT0:
            acc(b, t._emit(rt.OP_SELECT_GT0, a, g, zero)); acc(c, t._emit(rt.OP_SELECT_GT0, a, zero, g))
        elif op == rt.OP_ATAN2:               # d atan2(y, x) = (x dy - y dx) / (x^2 + y^2)
            w = t.div(g, t.add(t.mul(a, a), t.mul(b, b)))
            acc(a, t.mul(w, b)); acc(b, t.neg(t.mul(w, a)))
        else:
            raise NotImplementedError(f'derivative of opcode {op}')
    wanted = range(len(in_regs)) if wanted is None else wanted      # only these inputs' cotangents are produced
    outputs = [[adj.get(r, zero) for r in in_regs[k]] for k in wanted]
    return t.finish(outputs)
