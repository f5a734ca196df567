"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, on top of oracle/numga_oracle.py) of the reference's rigid-body XPBD SUB-STEP:
`Body.integrate(dt, constraint_sets)` of numga/examples/physics/core.py with the classes of
numga/examples/physics/base.py and `RK4` of numga/examples/integrators.py:7-12.  It takes the state arrays the
reference's `setup_bodies` produces (tests/golden/physics.npz `init/*`, `set{k}/*`, made by the unmodified reference
through tests/golden/make_golden_physics.py) and steps them.

Parity is PINNED: tests/test_oracle.py checks every stage of the first sub-step and the state after 1, 2 and 5
sub-steps against tests/golden/physics.npz (fp64 and fp32).

Only tests/, `__graft_entry__.smoke()` and bench.py's CPU legs (cpu_baseline of config 5, `--impl reference`) may
import this module; nothing under numga_b200/ does.
"""
import numpy as np

from oracle import numga_oracle as O


class DenseOp:
    """Unary operator with a float kernel [..., n_in, n_out]: what `partial` / `inverse()` / `[idx]` give
    (numga/backend/numpy/operator.py:56-70,78-90,102-132)."""

    def __init__(self, alg, kernel, in_blades, out_blades):
        self.alg, self.kernel, self.in_blades, self.out_blades = alg, kernel, tuple(in_blades), tuple(out_blades)

    def __getitem__(self, idx):
        return DenseOp(self.alg, self.kernel[idx], self.in_blades, self.out_blades)

    def __call__(self, mv):
        # NumpyEinsumOperator.__call__ (backend/numpy/operator.py:117-132): ONE einsum, optimize=True, into a
        # zero-initialised order='F' output of the broadcast shape
        assert mv.blades == self.in_blades
        shape = np.broadcast_shapes(self.kernel.shape[:-2], mv.values.shape[:-1]) + (len(self.out_blades),)
        out = np.zeros(shape, dtype=mv.values.dtype, order='F')
        np.einsum('...ab,...a->...b', self.kernel, mv.values, out=out, optimize=True)
        return O.MV(self.alg, self.out_blades, out)


def inertia_map(points):
    """`points.inertia_map()` (multivector/multivector.py:388-395): the ternary `inertia` operator with both point
    slots bound; kernel[..., bivector, antibivector] by ONE einsum (backend/numpy/operator.py:56-70)."""
    alg = points.alg
    biv = alg.subspace('bivector')
    op = O.inertia(alg, points.blades, biv)
    kernel = np.einsum('abcd,...a,...b->...cd', op.kernel, points.values, points.values, optimize=True)
    return DenseOp(alg, kernel, biv, op.out)


def RK4(f, y, h):                                                   # examples/integrators.py:7-12
    k1 = f(y)
    k2 = f(y + 0.5 * h * k1)
    k3 = f(y + 0.5 * h * k2)
    k4 = f(y + h * k3)
    return y + (h / 3) * (k2 + k3 + (k1 + k4) * 0.5)


def exp_linear_normalized(b):                                       # extension/logexp.py:88-92
    return (b + 1).normalized()


def motor_log_linear_normalized(m):                                 # extension/logexp.py:79-86
    return m._log_linear_normalized()


def motor_add_step(motor, step):                                    # core.py:33-35
    return motor * exp_linear_normalized(step * (-1 / 2))


def motor_relative_step(old, new):                                  # core.py:36-38
    return motor_log_linear_normalized(new.reverse_product(old)) * -(2)


def _take(mv, idx):
    return O.MV(mv.alg, mv.blades, mv.values[idx])


def _set(mv, idx, new):                                             # helper.py:19-23 + numpy context set_array
    v = mv.values.copy()
    v[idx] = new.values
    return O.MV(mv.alg, mv.blades, v)


class Body:
    """examples/physics/base.py:24-130 (state) + core.py:41-88 (dynamics)."""

    def __init__(self, motor, rate, first_moment, inertia, inertia_inv, damping, gravity):
        self.motor, self.rate, self.first_moment = motor, rate, first_moment
        self.inertia, self.inertia_inv, self.damping, self.gravity = inertia, inertia_inv, damping, gravity

    def copy(self, motor=None, rate=None):
        return Body(self.motor if motor is None else motor, self.rate if rate is None else rate, self.first_moment,
                    self.inertia, self.inertia_inv, self.damping, self.gravity)

    def forques(self):                                              # core.py:42-49
        damping = -(self.rate * self.damping).dual()
        gravity = self.first_moment & (self.motor << self.gravity)
        return damping + gravity

    def rate_derivative(self):                                      # core.py:51-57
        return self.inertia_inv(self.forques() - self.inertia(self.rate).commutator(self.rate))

    def pre_integrate(self, dt):                                    # core.py:68-72
        rate = RK4(lambda r: self.copy(rate=r).rate_derivative(), self.rate, dt)
        return self.copy(motor=motor_add_step(self.motor, rate * dt), rate=rate)

    def post_integrate(self, old, dt):                              # core.py:73-79
        motor = self.motor.normalized()
        rate = motor_relative_step(motor.reverse(), old.motor.reverse()) / dt
        return self.copy(rate=rate, motor=motor)

    def integrate(self, dt, constraint_sets=()):                    # core.py:80-88
        new = self.pre_integrate(dt)
        for c in constraint_sets:
            new = c.apply(new, dt=dt)
        new = new.post_integrate(self, dt=dt)
        for c in constraint_sets:
            new = c.v_apply(new, dt=dt)
        return new


class Constraint:
    """examples/physics/base.py:133-165 + core.py:90-167."""

    def __init__(self, body_idx, anchors, compliance, anchors_map=None):
        self.body_idx = np.asarray(body_idx, dtype=np.int64)
        self.anchors, self.compliance = anchors, compliance
        self.anchors_map = inertia_map(anchors) if anchors_map is None else anchors_map
        dtype = anchors.values.dtype
        self.connectivity = O.MV(anchors.alg, (0,), np.array([[+1 / 2], [-1 / 2]], dtype=dtype)[:, None])   # (2, 1, 1)

    def apply(self, bodies, dt):                                    # core.py:97-102
        idx = self.body_idx
        motors = self.apply_indexed(_take(bodies.motor, idx), bodies.inertia_inv[idx], dt)
        return bodies.copy(motor=_set(bodies.motor, idx, motors))

    def apply_indexed(self, motors, inertia_inv, dt):               # core.py:104-120
        anchors = motors >> self.anchors
        forque = _take(anchors, 0) & _take(anchors, 1)
        magnitude = forque.norm()
        direction = forque / (magnitude + 1e-26)
        steps = self.distribute_forque(motors << direction, magnitude, inertia_inv, self.compliance / dt ** 2)
        return motor_add_step(motors, steps)

    def v_apply(self, bodies, dt):                                  # core.py:122-127
        idx = self.body_idx
        rates = self.v_apply_indexed(_take(bodies.motor, idx), _take(bodies.rate, idx), bodies.inertia_inv[idx], dt)
        return bodies.copy(rate=_set(bodies.rate, idx, rates))

    def v_apply_indexed(self, motors, rates, inertia_inv, dt):      # core.py:129-146
        velocities = motors >> self.anchors_map(rates)
        weighted = self.connectivity * velocities
        forque = -O.MV(weighted.alg, weighted.blades, weighted.values.sum(0))
        magnitude = forque.norm()
        direction = forque / (magnitude + 1e-26)
        steps = self.distribute_forque(motors << direction, magnitude, inertia_inv, self.compliance * 0)
        return rates + steps

    def distribute_forque(self, directions, magnitude, inertias_inv, constraint_compliance):   # core.py:148-167
        steps = inertias_inv(directions)
        inertial_compliances = steps & directions
        total = constraint_compliance + O.MV(steps.alg, inertial_compliances.blades, inertial_compliances.values.sum(0))
        multiplier = magnitude / total
        return steps * self.connectivity * multiplier


def from_arrays(arrays, dtype, n_chains=1, rate=None):
    """Bodies + constraint sets from the reference's set-up arrays (dict with the `init/*`, `set{k}/*` entries of
    tests/golden/physics.npz, prefix stripped), replicated to `n_chains` independent chains laid out chain-major
    (block-diagonal body_idx): the config-5 workload."""
    alg = O.Algebra((3, 0, 1))
    dtype = np.dtype(dtype)

    def rep(a, axis=0):
        a = np.asarray(a, dtype=dtype)
        reps = [1] * a.ndim
        reps[axis] = n_chains
        return np.asfortranarray(np.tile(a, reps)) if n_chains > 1 else np.asarray(a, dtype=dtype, order='F')
    biv, abiv = alg.subspace('bivector'), alg.order(O.right_hodge(alg, alg.subspace('bivector')).out)
    grav_blades = tuple(int(b) for b in arrays['init/gravity_blades'])
    nb = arrays['init/motor'].shape[0]
    bodies = Body(
        O.MV(alg, alg.subspace('even_grade'), rep(arrays['init/motor'])),
        O.MV(alg, biv, rep(arrays['init/rate']) if rate is None else
             (np.asarray(rate, dtype=dtype, order='F') if len(rate) == nb * n_chains else rep(rate))),
        O.MV(alg, alg.subspace('antivector'), rep(arrays['init/first_moment'])),
        DenseOp(alg, rep(arrays['init/inertia']), biv, abiv),
        DenseOp(alg, rep(arrays['init/inertia_inv']), abiv, biv),
        O.MV(alg, (0,), rep(arrays['init/damping'])),
        O.MV(alg, grav_blades, rep(arrays['init/gravity'])))
    sets = []
    k = 0
    while f'set{k}/body_idx' in arrays:
        idx = np.asarray(arrays[f'set{k}/body_idx'], dtype=np.int64)
        nc = idx.shape[1]
        if n_chains > 1:
            idx = np.tile(idx, (1, n_chains)) + np.repeat(np.arange(n_chains) * nb, nc)[None, :]
        anchors = O.MV(alg, alg.subspace('antivector'), rep(arrays[f'set{k}/anchors'], axis=1))
        amap = DenseOp(alg, rep(arrays[f'set{k}/anchors_map'], axis=1), biv, abiv)
        sets.append(Constraint(idx, anchors, O.MV(alg, (0,), rep(arrays[f'set{k}/compliance'])), amap))
        k += 1
    return bodies, sets
