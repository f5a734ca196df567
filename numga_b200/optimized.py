"""Opt-in closed forms for 3D PGA (signature +++0): `normalized` of motors, `exp` of bivectors, `motor_log` of motors.

Mirrors the reference's OPTIONAL `numga/multivector/extension/optimized.py` (not imported or active by default there
either): the default dispatch keeps the generic, signature-agnostic algorithms (bisection exp: ~570 instructions per
item, compute-bound), these closed forms (one sqrt, one sin, one cos) make `exp` an HBM-bound kernel.  Same formulas
as the reference (optimized.py:24-50, 54-127; terms involving xz carry the sign of the alphabetic blade order), written
on traced values, so each still becomes ONE generated kernel and fuses with whatever surrounds it.

    import numga_b200.optimized as opt
    opt.enable()        # registers at position 0 of the `normalized` / `exp` dispatch, like the reference
    opt.disable()       # removes them again

`motor_log` has no closed form in the reference (its log is the 16-square-root bisection, logexp.py:101-117); the one
here is the exact inverse of the reference's closed-form `exp` above (principal branch, rotation angle in [0, pi]), so
that the whole exp -> normalized -> motor_log round trip of BASELINE config 3 becomes ONE HBM-bound kernel when the
closed forms are enabled.  It is checked against the bisection logarithm in fp64 (tests/test_optimized.py).

One deliberate difference: the reference's full-bivector `exp` computes `t = m (c - s) / l` and returns nan for a
bivector without rotation part (l = 0); here that limit is taken (t = -m / 3 as l -> 0).
"""
import numpy as np

from numga_b200 import runtime as rt
from numga_b200.extension import exp, motor_log, normalized
from numga_b200.trace import SymArray

_MOTOR = (0, 3, 5, 6, 9, 10, 12, 15)          # 1, xy, xz, yz, xw, yw, zw, xyzw
_BIVECTOR = (3, 5, 6, 9, 10, 12)
_ROT = (3, 5, 6)
_TRANS = (9, 10, 12)


def _is_3dpga(s, blades):
    d = s.algebra.description
    return d.signature_str == '+++0' and tuple(int(b) for b in s.blades) == blades


def _parts(values):
    return [values[..., i] for i in range(len(values))]


def _pack(context, subspace, parts):
    trace = parts[0].trace
    return context.multivector(values=SymArray(trace, [p.regs[0] for p in parts]), subspace=subspace)


def _unary(x, op):
    return x.map(lambda r: x.trace.unary(op, r))


def _sinc_cos(l):
    """a = sqrt(l): (cos a, sin a / a with the a -> 0 limit 1), optimized.py:75-79."""
    a = l ** 0.5
    c, sn = _unary(a, rt.OP_COS), _unary(a, rt.OP_SIN)
    t = a.trace
    eps, one = t.const(1e-20), t.const(1.0)
    s = SymArray(t, [t.select_gt0(t.sub(a.regs[0], eps), t.div(sn.regs[0], a.regs[0]), one)])
    return a, c, s


def normalize_3dpga_motor(m):
    """optimized.py:24-50."""
    e, xy, xz, yz, xw, yw, zw, xyzw = _parts(m.values)
    s = (e * e + xy * xy + xz * xz + yz * yz) ** (-0.5)
    d = (e * xyzw - xy * zw + xz * yw - yz * xw) * s * s
    out = [e * s, xy * s, xz * s, yz * s, xw * s + yz * s * d, yw * s - xz * s * d, zw * s + xy * s * d, xyzw * s - e * s * d]
    return _pack(m.context, m.subspace, out)


def exp_degenerate_bivector(b):
    """exp of a pure translation generator: 1 + b exactly (optimized.py:54-66)."""
    xw, yw, zw = _parts(b.values)
    one = SymArray(xw.trace, [xw.trace.const(1.0)])
    return _pack(b.context, b.context.subspace.translator(), [one, xw, yw, zw])


def exp_nondegenerate_bivector(b):
    """Rotor from a rotation generator (optimized.py:68-90)."""
    xy, xz, yz = _parts(b.values)
    _, c, s = _sinc_cos(xy * xy + xz * xz + yz * yz)
    return _pack(b.context, b.context.subspace.rotor(), [c, s * xy, s * xz, s * yz])


def exp_bivector(b):
    """Motor from a general bivector (optimized.py:92-127)."""
    xy, xz, yz, xw, yw, zw = _parts(b.values)
    l = xy * xy + xz * xz + yz * yz
    a, c, s = _sinc_cos(l)
    m = xy * zw - xz * yw + yz * xw
    t_ = a.trace
    # t = m (c - s) / l, with its l -> 0 limit -m / 3 (the reference divides by zero there)
    ratio = t_.select_gt0(t_.sub(l.regs[0], t_.const(1e-30)), t_.div(t_.sub(c.regs[0], s.regs[0]), l.regs[0]), t_.const(-1.0 / 3.0))
    t = m * SymArray(t_, [ratio])
    out = [c, s * xy, s * xz, s * yz, s * xw + t * yz, s * yw - t * xz, s * zw + t * xy, m * s]
    return _pack(b.context, b.context.subspace.motor(), out)


def log_motor(M):
    """Bivector b with exp(b) = M for a NORMALIZED 3D PGA motor: the inverse of `exp_bivector`.
    With L = XY^2 + XZ^2 + YZ^2 = sin^2 a:  a = atan2(sqrt L, e);  1/s = a / sin a;  m = P / s;
    t = m (cos a - s) / a^2;  rotation part = (XY, XZ, YZ) / s;  translation part = ((XW, YW, ZW) -+ t (yz, xz, xy)) / s."""
    e, XY, XZ, YZ, XW, YW, ZW, P = _parts(M.values)
    t_ = e.trace
    L = XY * XY + XZ * XZ + YZ * YZ
    sL = L ** 0.5
    a = SymArray(t_, [t_.atan2(sL.regs[0], e.regs[0])])
    # 1 / s = a / sin a, -> 1 as the rotation vanishes
    rs = SymArray(t_, [t_.select_gt0(t_.sub(sL.regs[0], t_.const(1e-20)), t_.div(a.regs[0], sL.regs[0]), t_.const(1.0))])
    xy, xz, yz = XY * rs, XZ * rs, YZ * rs
    m = P * rs
    l = a * a
    s = SymArray(t_, [t_.select_gt0(t_.sub(sL.regs[0], t_.const(1e-20)), t_.div(sL.regs[0], a.regs[0]), t_.const(1.0))])
    # (cos a - sin a / a) / a^2 = -1/3 + a^2/30 - a^4/840 + ...: the series where the difference would cancel
    series = t_.add(t_.const(-1.0 / 3.0), t_.mul(l.regs[0], t_.sub(t_.const(1.0 / 30.0), t_.mul(l.regs[0], t_.const(1.0 / 840.0)))))
    ratio = t_.select_gt0(t_.sub(l.regs[0], t_.const(0.01)), t_.div(t_.sub(e.regs[0], s.regs[0]), l.regs[0]), series)
    t = m * SymArray(t_, [ratio])
    out = [xy, xz, yz, (XW - t * yz) * rs, (YW + t * xz) * rs, (ZW - t * xy) * rs]
    return _pack(M.context, M.context.subspace.bivector(), out)


_REGISTERED = []


def enable():
    """Register the closed forms in front of the generic implementations (idempotent)."""
    if _REGISTERED:
        return
    for dispatch, cond, fn in (
            (normalized, lambda s: _is_3dpga(s, _MOTOR), normalize_3dpga_motor),
            (exp, lambda s: _is_3dpga(s, _TRANS), exp_degenerate_bivector),
            (exp, lambda s: _is_3dpga(s, _ROT), exp_nondegenerate_bivector),
            (exp, lambda s: _is_3dpga(s, _BIVECTOR), exp_bivector),
            (motor_log, lambda s: _is_3dpga(s, _MOTOR), log_motor)):
        dispatch.register(cond, position=0)(fn)
        _REGISTERED.append((dispatch, fn))
    _invalidate()


def disable():
    for dispatch, fn in _REGISTERED:
        dispatch.options[:] = [(c, f) for c, f in dispatch.options if f is not fn]
        dispatch.cache.clear()
    _REGISTERED.clear()
    _invalidate()


def _invalidate():
    """Compiled-function caches are keyed on the method name, not on the dispatch table: drop them in every live
    context (the reference does the same with `mv.exp.cache = {}`, optimized.py:52,129)."""
    from numga_b200.backend import B200Context
    for ctx in list(B200Context.instances or ()):          # (None until the first context exists)
        ctx.engine.cache.clear()
        ctx.engine.fast_cache.clear()
        ctx.lazy_engine.signatures.clear()
