"""XPBD rigid-body chain in geometric algebra — the core step of numga's physics example, on the B200 backend.

Mirrors `numga/examples/physics/{core,base,setup_chain}.py` and `examples/integrators.py` (same class and method
names, same formulas: Verlet-style integrate = pre_integrate (RK4 on Euler's equation) -> constraint relaxation
per colour set -> post_integrate -> velocity relaxation per colour set).  Two ways to run a step:

  * `Body.integrate(dt, constraint_sets)`        generic path, written like the reference on `[2, c]`-shaped
    multivectors; every multivector method is one generated kernel (174 reference array passes -> ~120 kernels);
  * `FusedStep(bodies, constraint_sets).integrate(bodies, dt)`   the B200 path: SIX kernels per sub-step for two
    colour sets — pre-integrate, constraint-apply x2, post-integrate, velocity-apply x2 — each traced from the
    same method bodies below with `context.jit`.  The constraint kernels GATHER the two bodies of every constraint
    through an index operand and SCATTER the updated motors / rates back in place (each body is touched at most
    once per colour set, numga/examples/physics/setup_chain.py:66-67, so the scatter is race-free).
"""
import numpy as np

from numga_b200.backend import B200DenseOperator


def RK4(f, y, h):
    """Classic Runge-Kutta step (examples/integrators.py:7-12)."""
    k1 = f(y)
    k2 = f(y + 0.5 * h * k1)
    k3 = f(y + 0.5 * h * k2)
    k4 = f(y + h * k3)
    return y + (h / 3) * (k2 + k3 + (k1 + k4) * 0.5)


def motor_add_step(motor, step):
    """Apply an integrated-rate bivector to a motor; linearised exp (core.py:33-35)."""
    return motor * (step * (-1 / 2)).exp_linear_normalized()


def motor_relative_step(old, new):
    """Bivector step between two motor states (core.py:36-38)."""
    return new.reverse_product(old).motor_log_linear_normalized() * -(2)


class Body:
    """A set of rigid bodies (leading batch axis).  motor: body->world; rate: bivector, body frame;
    first_moment: sum of mass points; inertia / inertia_inv: bound unary operators (rates <-> momentum lines)."""

    def __init__(self, motor, rate, first_moment, inertia, inertia_inv, damping, gravity):
        assert motor.subspace.equals.motor()
        assert rate.subspace.equals.bivector()
        assert gravity.subspace.inside.antivector()
        assert inertia.output.equals.antibivector()
        assert inertia_inv.output.equals.bivector()
        assert damping.subspace.equals.scalar()
        self.motor, self.rate, self.first_moment = motor, rate, first_moment
        self.inertia, self.inertia_inv, self.damping, self.gravity = inertia, inertia_inv, damping, gravity

    _FIELDS = ('motor', 'rate', 'first_moment', 'inertia', 'inertia_inv', 'damping', 'gravity')

    def __getitem__(self, idx):
        return type(self)(*(getattr(self, f)[idx] for f in self._FIELDS))

    def concatenate(self, other):
        return type(self)(*(getattr(self, f).concatenate(getattr(other, f), axis=0) for f in self._FIELDS))

    def copy(self, motor=None, rate=None, gravity=None, damping=None):
        return type(self)(
            motor=self.motor if motor is None else motor, rate=self.rate if rate is None else rate,
            damping=self.damping if damping is None else damping, gravity=self.gravity if gravity is None else gravity,
            first_moment=self.first_moment, inertia=self.inertia, inertia_inv=self.inertia_inv)

    @classmethod
    def from_point_cloud(cls, points):
        """Body from mass points [..., n_points] (base.py:92-124): first moment = sum of points, inertia map =
        sum of the per-point maps  rate -> point & (point x rate)."""
        context = points.context
        assert points.subspace.inside.antivector()
        first_moment = points.sum(axis=-2)
        inertia = points.inertia_map().sum(axis=-3)
        ones = np.ones(first_moment.shape)
        return cls(
            motor=context.multivector.motor() * ones, rate=context.multivector.bivector() * ones,
            first_moment=first_moment, inertia=inertia, inertia_inv=inertia.inverse(),
            damping=context.multivector.scalar() * ones * 0, gravity=context.multivector.antivector() * ones)

    def kinetic_energy(self):
        return self.inertia(self.rate) & self.rate

    # ---- dynamics (core.py:41-88) -------------------------------------------------------------------------
    def forques(self):
        """External forque line per body, body frame: linear damping + gravity acting on the first moment."""
        damping = -(self.rate * self.damping).dual()
        gravity = self.first_moment & (self.motor << self.gravity)
        return damping + gravity

    def rate_derivative(self):
        """Euler's equation: d/dt rate = I^-1 (forque - I(rate) x rate)."""
        return self.inertia_inv(self.forques() - self.inertia(self.rate).commutator(self.rate))

    def pre_integrate(self, dt):
        rate = RK4(lambda r: self.copy(rate=r).rate_derivative(), self.rate, dt)
        return self.copy(motor=motor_add_step(self.motor, rate * dt), rate=rate)

    def post_integrate(self, old, dt):
        motor = self.motor.normalized()
        rate = motor_relative_step(motor.reverse(), old.motor.reverse()) / dt
        return self.copy(rate=rate, motor=motor)

    def integrate(self, dt, constraint_sets=()):
        new = self.pre_integrate(dt)
        for c in constraint_sets:
            new = c.apply(new, dt=dt)
        new = new.post_integrate(self, dt=dt)
        for c in constraint_sets:
            new = c.v_apply(new, dt=dt)
        return new


class Constraint:
    """Point-to-point constraints: anchors [2, n] are the body-local attachment points of bodies body_idx [2, n]."""

    def __init__(self, body_idx, anchors, compliance):
        context = anchors.context
        assert anchors.subspace.equals.antivector()
        assert compliance.subspace.equals.scalar()
        self.body_idx = context.coerce_array(body_idx, dtype=int)
        self.anchors = anchors
        self.anchors_map = anchors.inertia_map()          # rates -> velocity line at the anchor
        self.compliance = compliance
        self.connectivity = context.multivector.scalar([[+1 / 2], [-1 / 2]])[:, None]   # +-1/2 per side, shape (2, 1)
        assert self.connectivity.shape == (2, 1)

    def __getitem__(self, idx):
        return type(self)(self.body_idx[:, idx], self.anchors[:, idx], self.compliance[idx])

    def concatenate(self, other):
        import torch
        return type(self)(torch.cat([self.body_idx, other.body_idx], dim=1),
                          self.anchors.concatenate(other.anchors, axis=1),
                          self.compliance.concatenate(other.compliance, axis=0))

    def apply(self, bodies, dt):
        idx = self.body_idx
        motors = self.apply_indexed(bodies.motor[idx], bodies.inertia_inv[idx], dt)
        return bodies.copy(motor=bodies.motor.at[idx].set(motors))

    def apply_indexed(self, motors, inertia_inv, dt):
        """Move both bodies so that the anchors meet, balancing constraint work and inertial work (core.py:104-120)."""
        anchors = motors >> self.anchors                     # [2, c] anchors in world space
        forque = anchors[0] & anchors[1]                     # [c] connecting line
        magnitude = forque.norm()                            # [c] violation
        direction = forque / (magnitude + 1e-26)
        steps = self.distribute_forque(motors << direction, magnitude, inertia_inv, self.compliance / dt ** 2)
        return motor_add_step(motors, steps)

    def v_apply(self, bodies, dt):
        idx = self.body_idx
        rates = self.v_apply_indexed(bodies.motor[idx], bodies.rate[idx], bodies.inertia_inv[idx], dt)
        return bodies.copy(rate=bodies.rate.at[idx].set(rates))

    def v_apply_indexed(self, motors, rates, inertia_inv, dt):
        """Exchange a velocity impulse along the constraint (core.py:129-146)."""
        velocities = motors >> self.anchors_map(rates)       # [2, c] world-space velocity lines
        forque = -(self.connectivity * velocities).sum(axis=0)
        magnitude = forque.norm()
        direction = forque / (magnitude + 1e-26)
        steps = self.distribute_forque(motors << direction, magnitude, inertia_inv, self.compliance * 0)
        return rates + steps

    def distribute_forque(self, directions, magnitude, inertias_inv, constraint_compliance):
        """Split an impulse along `directions` [2, c] into momentum-conserving bivector steps (core.py:148-167)."""
        steps = inertias_inv(directions)
        inertial_compliances = steps & directions
        total_compliance = constraint_compliance + inertial_compliances.sum(axis=0)
        multiplier = magnitude / total_compliance
        return steps * self.connectivity * multiplier


def setup_bodies(context, fixing=True, n_bodies=12, compliance=1e-9, distance=9e-2, size=5e-2, damping=1e-3):
    """A chain of identical bodies linked by point constraints, first body pinned (setup_chain.py:8-67).
    Returns (bodies, [red constraints, black constraints])."""
    *axes, origin = context.multivector.basis()            # last basis vector = origin / projective direction
    model_space = axes[0].subspace
    for a in axes[1:]:
        model_space = model_space.union(a.subspace)
    horizon = origin.subspace.wedge(model_space)           # translation-like bivector directions
    ndim = len(axes)

    def translator(d):
        return ((origin / -2) ^ context.multivector(model_space, d)).exp()

    def point_embed(d):
        return translator(d) >> origin.dual_inverse()

    cube = np.array(np.kron(np.diag(np.arange(ndim) + 1) / ndim, [+1, -1]).T)     # 2*ndim mass points
    bodies = Body.from_point_cloud(points=point_embed(cube * size))
    bodies = bodies.copy(gravity=(axes[0] * 5e-2).dual_inverse())
    bodies = bodies.copy(damping=bodies.damping + damping)

    d = np.arange(n_bodies) * distance
    qs = translator(np.array([d * 0] * (ndim - 1) + [d]).T)
    bodies = bodies[None][np.zeros(n_bodies, int)]
    bodies = bodies.copy(motor=bodies.motor * qs)

    i = np.arange(n_bodies - 1)
    body_idx = np.array([i, i + 1])
    a = np.array([[0] * (ndim - 1) + [distance], [0] * (ndim - 1) + [-distance]]) / 2
    anchors = point_embed(a[:, None, :] * np.ones((1, n_bodies - 1, 1)))
    ones = context.multivector.scalar(np.ones((n_bodies - 1, 1)))
    constraints = Constraint(body_idx, anchors, compliance=ones * compliance)
    if fixing:      # pin body 0: zero the translation-like columns of its inverse inertia
        idx = bodies.inertia_inv.output.to_relative_indices(horizon.blades)
        bodies.inertia_inv = bodies.inertia_inv.at[0, :, idx].set(0)
    return bodies, [constraints[0::2], constraints[1::2]]


def replicate_chains(bodies, constraint_sets, n_chains):
    """n_chains independent copies of a chain laid out chain-major (block-diagonal body_idx): the config-5
    workload.  Chains never interact, so the batch shards across GPUs by chain with no exchange."""
    import torch
    nb = bodies.motor.shape[0]
    rep = lambda mv: mv[np.tile(np.arange(mv.shape[0]), n_chains)]
    big = Body(*(rep(getattr(bodies, f)) for f in Body._FIELDS))
    sets = []
    for c in constraint_sets:
        nc = c.body_idx.shape[1]
        offs = (torch.arange(n_chains, device=c.body_idx.device) * nb).repeat_interleave(nc)
        sets.append(Constraint(c.body_idx.repeat(1, n_chains) + offs[None, :], c.anchors[:, np.tile(np.arange(nc), n_chains)],
                               c.compliance[np.tile(np.arange(nc), n_chains)]))
    return big, sets


class FusedStep:
    """The sub-step as six generated kernels (see module docstring).  State arrays are updated IN PLACE on buffers
    this object owns, so one `integrate` reads and writes each state component a small fixed number of times."""

    def __init__(self, bodies, constraint_sets):
        ctx = self.context = bodies.motor.context
        self.sets = constraint_sets
        jit = ctx.jit

        def pre(motor, rate, first_moment, inertia, inertia_inv, damping, gravity, dt):
            b = Body(motor, rate, first_moment, inertia, inertia_inv, damping, gravity)
            new = b.pre_integrate(dt)
            return new.motor, new.rate
        self._pre = jit(pre)

        def post(motor, old_motor, dt):
            m = motor.normalized()
            return m, motor_relative_step(m.reverse(), old_motor.reverse()) / dt
        self._post = jit(post)

        def apply(m0, m1, inv0, inv1, a0, a1, compliance, dt):
            w0, w1 = m0 >> a0, m1 >> a1
            forque = w0 & w1
            magnitude = forque.norm()
            direction = forque / (magnitude + 1e-26)
            s0, s1 = self._distribute(m0 << direction, m1 << direction, magnitude, inv0, inv1, compliance / (dt * dt))
            return motor_add_step(m0, s0), motor_add_step(m1, s1)
        self._apply = jit(apply)

        def v_apply(m0, m1, r0, r1, inv0, inv1, map0, map1, compliance, dt):
            v0, v1 = m0 >> map0(r0), m1 >> map1(r1)
            forque = -(v0 * 0.5 + v1 * -0.5)
            magnitude = forque.norm()
            direction = forque / (magnitude + 1e-26)
            s0, s1 = self._distribute(m0 << direction, m1 << direction, magnitude, inv0, inv1, compliance * 0)
            return r0 + s0, r1 + s1
        self._v_apply = jit(v_apply)

    @staticmethod
    def _distribute(d0, d1, magnitude, inv0, inv1, constraint_compliance):
        s0, s1 = inv0(d0), inv1(d1)
        total = constraint_compliance + ((s0 & d0) + (s1 & d1))
        multiplier = magnitude / total
        return s0 * 0.5 * multiplier, s1 * -0.5 * multiplier

    def integrate(self, bodies, dt):
        motor, rate = self._pre(bodies.motor, bodies.rate, bodies.first_moment, bodies.inertia, bodies.inertia_inv,
                                bodies.damping, bodies.gravity, dt)
        for c in self.sets:
            i0, i1 = c.body_idx[0], c.body_idx[1]
            self._apply(motor.indexed(i0), motor.indexed(i1), bodies.inertia_inv.indexed(i0),
                        bodies.inertia_inv.indexed(i1), c.anchors[0], c.anchors[1], c.compliance, dt,
                        out=(motor.indexed(i0), motor.indexed(i1)))
        motor, rate = self._post(motor, bodies.motor, dt)
        for c in self.sets:
            i0, i1 = c.body_idx[0], c.body_idx[1]
            self._v_apply(motor.indexed(i0), motor.indexed(i1), rate.indexed(i0), rate.indexed(i1),
                          bodies.inertia_inv.indexed(i0), bodies.inertia_inv.indexed(i1), c.anchors_map[0],
                          c.anchors_map[1], c.compliance, dt, out=(rate.indexed(i0), rate.indexed(i1)))
        return bodies.copy(motor=motor, rate=rate)


def _block_structure(n_bodies, index_sets, max_block=1024):
    """Smallest block size b such that the bodies split into n_bodies / b consecutive blocks, every constraint joins
    two bodies of ONE block, and the constraints of every set are laid out block by block with the same count per
    block.  Returns (b, [constraints per block per set]); raises when the system has no such structure."""
    import torch
    for b in range(2, min(max_block, n_bodies) + 1):
        if n_bodies % b:
            continue
        n_blocks = n_bodies // b
        per_block, ok = [], True
        for idx in index_sets:
            nc = int(idx.shape[1])
            if nc % n_blocks:
                ok = False
                break
            cpb = nc // n_blocks
            blk = torch.arange(nc, device=idx.device) // cpb
            if not bool((torch.div(idx, b, rounding_mode='floor') == blk[None, :]).all()):
                ok = False
                break
            per_block.append(cpb)
        if ok:
            return b, per_block
    raise ValueError('constraints do not decompose into independent blocks of <= %d consecutive bodies; use FusedStep' % max_block)


def _partner(x):
    """`x` as computed by the thread of the OTHER side of the same constraint (items 2k and 2k + 1 are the two sides:
    adjacent lanes, one warp shuffle per component)."""
    from numga_b200.trace import SymArray
    v = x.values
    return x.context.multivector(values=SymArray(v.trace, [v.trace.partner(r) for r in v.regs]), subspace=x.subspace)


def _ordered(side, mine, theirs):
    """(value of side 0, value of side 1) given this thread's and its partner's value; `side` > 0 on side 0."""
    from numga_b200.trace import SymArray
    t, s = mine.values.trace, side.values.regs[0]
    pick = lambda a, b: mine.context.multivector(
        values=SymArray(t, [t.select_gt0(s, ra, rb) for ra, rb in zip(a.values.regs, b.values.regs)]), subspace=mine.subspace)
    return pick(mine, theirs), pick(theirs, mine)


class ChainStep:
    """The sub-step as ONE chain-resident kernel (numga_b200/pipeline.py): a CTA owns a tile of whole chains, loads
    motor / rate / inverse inertia into shared memory once, runs  pre -> apply per colour set -> post -> v_apply per
    colour set  with CTA barriers in between (constraint threads read and write the two bodies they join in shared
    memory), repeats that `unroll` times (the reference drives `unroll=10` sub-steps per frame,
    numga/examples/physics/run_chain_numpy.py:26-30) and stores motor / rate once.  Per launch every state word
    crosses HBM once; the per-body constants and per-constraint tables are re-read per sub-step from L2.

    Needs the block structure `replicate_chains` produces (independent chains laid out one after the other); the
    formulas are the same traced python bodies `FusedStep` uses.  `integrate(bodies, dt)` = `unroll` sub-steps of dt."""

    def __init__(self, bodies, constraint_sets, unroll=1, bodies_per_tile=None, threads_per_tile=0, exploit_zeros=False,
                 recompute_maps=None, pairs=False):
        from numga_b200.pipeline import Pipeline, compact_operator, expand_operator
        from numga_b200.backend import _OperatorKey, kernel_space
        import torch
        ctx = self.context = bodies.motor.context
        self.sets, self.unroll = list(constraint_sets), int(unroll)
        n_bodies = bodies.motor.shape[0]
        block, per_block = _block_structure(n_bodies, [c.body_idx for c in self.sets])
        if bodies_per_tile is None:
            bodies_per_tile = 384 if ctx.dtype == torch.float32 else 192      # measured on B200 (profiles/r02_chain_sweep.txt)
        k = max(1, bodies_per_tile // block)                       # whole blocks (chains) per tile
        self.block, self.blocks_per_tile = block, k
        # `pairs`: TWO threads per constraint in the constraint phases, one per side -- the reference's own formulation (its
        # constraint arrays are [2, c]: numga/examples/physics/core.py:104-167) with the two values that couple the sides
        # (the other anchor / velocity line, the other side's inertial compliance) exchanged by warp shuffle
        # (NGB200_OP_PARTNER).  Half the registers per thread, twice the threads: one body per thread in the body phases.
        self.pairs = bool(pairs)
        if self.pairs:
            assert not exploit_zeros, 'pairs=True re-evaluates the anchor maps (no tables to inspect)'
            per_block = [2 * c for c in per_block]
        if not threads_per_tile:
            # one thread per constraint in the constraint phases (every warp busy there: they are 60 % of the work), two
            # bodies per thread in the body phases: 8.07 vs 6.60 G body-steps/s against one thread per body
            threads_per_tile = min(1024, -(-max([c * k for c in per_block] + [-(-block * k // 2)]) // 32) * 32)
        # Which entries of the bound operators (inertia, inverse inertia, anchor maps) can be non-zero.  By default that is
        # what the integer tables say (`op.mask`; an `inverse()` is dense).  With `exploit_zeros` the tables' VALUES are
        # inspected once, here: entries that are exactly zero for every item are compiled out (never loaded, never
        # multiplied) -- the "numerical sparsity of inertia tensors" the reference notes but cannot use
        # (numga/backend/torch/operator.py:60-64).  The kernel is then specific to this zero pattern; `integrate` checks
        # that it is handed the same (unmodified) tables.
        self.exploit_zeros = bool(exploit_zeros)

        def mask_of(op):
            if not self.exploit_zeros or ctx.compile_only:        # (compile-only contexts hold uninitialised tables)
                return op.mask
            flat = op.flat.reshape(-1, op.flat.shape[-1])
            return op.mask & (flat != 0).any(dim=0).cpu().numpy().reshape(op.kshape)
        inv_mask, inertia_mask = mask_of(bodies.inertia_inv), mask_of(bodies.inertia)
        map_masks = [mask_of(c.anchors_map) for c in self.sets]
        self._inspected = [(t.data_ptr(), t._version) for t in (bodies.inertia.flat, bodies.inertia_inv.flat)]
        inv_axes = bodies.inertia_inv.axes
        S = ctx.subspace
        names = ['body'] + [f'set{j}' for j in range(len(self.sets))]
        pipe = Pipeline(ctx, dict(zip(names, [block * k] + [c * k for c in per_block])), threads_per_tile=threads_per_tile)
        inv_space = _OperatorKey(inv_axes, inv_mask)
        inv_compact = kernel_space((int(inv_mask.sum()),))          # shared memory holds the non-zero entries only
        inertia_space = _OperatorKey(bodies.inertia.axes, inertia_mask)
        g_motor, g_rate = pipe.array('motor', S.even_grade(), 'body'), pipe.array('rate', S.bivector(), 'body')
        g_fm = pipe.array('first_moment', bodies.first_moment.subspace, 'body')
        g_inertia, g_inv = pipe.array('inertia', inertia_space, 'body'), pipe.array('inertia_inv', inv_space, 'body')
        g_damping, g_gravity = pipe.array('damping', bodies.damping.subspace, 'body'), pipe.array('gravity', bodies.gravity.subspace, 'body')
        o_motor, o_rate = pipe.array('motor_out', S.even_grade(), 'body'), pipe.array('rate_out', S.bivector(), 'body')
        s_motor, s_rate = pipe.buffer(S.even_grade(), 'body'), pipe.buffer(S.bivector(), 'body')
        s_old, s_inv = pipe.buffer(S.even_grade(), 'body'), pipe.buffer(inv_compact, 'body')
        dt = pipe.param(0)
        unpack = lambda c: expand_operator(ctx, inv_axes, inv_mask, c)

        pipe.phase('body', lambda m, r, inv: (m, r, compact_operator(inv, inv_mask)), [g_motor, g_rate, g_inv], [s_motor, s_rate, s_inv])

        def pre(motor, rate, first_moment, inertia, inv_c, damping, gravity, dt):
            new = Body(motor, rate, first_moment, inertia, unpack(inv_c), damping, gravity).pre_integrate(dt)
            return new.motor, new.rate, motor
        pipe.phase('body', pre, [s_motor, s_rate, g_fm, g_inertia, s_inv, g_damping, g_gravity, dt], [s_motor, s_rate, s_old])

        def apply(m0, m1, inv0, inv1, a0, a1, compliance, dt):
            inv0, inv1 = unpack(inv0), unpack(inv1)
            w0, w1 = m0 >> a0, m1 >> a1
            forque = w0 & w1
            magnitude = forque.norm()
            direction = forque / (magnitude + 1e-26)
            s0, s1 = FusedStep._distribute(m0 << direction, m1 << direction, magnitude, inv0, inv1, compliance / (dt * dt))
            return motor_add_step(m0, s0), motor_add_step(m1, s1)

        def post(motor, old_motor, dt):
            m = motor.normalized()
            return m, motor_relative_step(m.reverse(), old_motor.reverse()) / dt

        def v_apply(m0, m1, r0, r1, inv0, inv1, map0, map1, compliance, dt):
            inv0, inv1 = unpack(inv0), unpack(inv1)
            v0, v1 = m0 >> map0(r0), m1 >> map1(r1)
            forque = -(v0 * 0.5 + v1 * -0.5)
            magnitude = forque.norm()
            direction = forque / (magnitude + 1e-26)
            s0, s1 = FusedStep._distribute(m0 << direction, m1 << direction, magnitude, inv0, inv1, compliance * 0)
            return r0 + s0, r1 + s1

        # `Constraint.anchors_map` (rates -> velocity line at the anchor, base.py:147) is `anchors.inertia_map()`: 24
        # quadratic forms of the 4 anchor coordinates.  The reference tabulates it once because every numpy pass is
        # expensive; here re-evaluating it in registers (~150 flops per constraint and sub-step) is cheaper than
        # streaming 2 x 24 table words per constraint through HBM / L2 every sub-step (29 % of the step's bytes).
        # The arithmetic is the one that produced the table, so results do not change.
        self.recompute_maps = (not self.exploit_zeros) if recompute_maps is None else bool(recompute_maps)

        def v_apply_from_anchors(m0, m1, r0, r1, inv0, inv1, a0, a1, compliance, dt):
            return v_apply(m0, m1, r0, r1, inv0, inv1, a0.inertia_map(), a1.inertia_map(), compliance, dt)

        def share(side, mine):
            return _ordered(side, mine, _partner(mine))

        def apply_side(m, inv, a, compliance, side, dt):
            """One side of `Constraint.apply` (core.py:104-120): this thread's body, anchor and inverse inertia."""
            inv = unpack(inv)
            w0, w1 = share(side, m >> a)
            forque = w0 & w1
            magnitude = forque.norm()
            direction = forque / (magnitude + 1e-26)
            d = m << direction
            s = inv(d)
            c0, c1 = share(side, s & d)
            multiplier = magnitude / (compliance / (dt * dt) + (c0 + c1))
            return motor_add_step(m, s * (side * 0.5) * multiplier)

        def v_apply_side(m, r, inv, a, compliance, side, dt):
            """One side of `Constraint.v_apply` (core.py:129-146)."""
            inv = unpack(inv)
            v0, v1 = share(side, m >> a.inertia_map()(r))
            forque = -(v0 * 0.5 + v1 * -0.5)
            magnitude = forque.norm()
            direction = forque / (magnitude + 1e-26)
            d = m << direction
            s = inv(d)
            c0, c1 = share(side, s & d)
            multiplier = magnitude / (compliance * 0 + (c0 + c1))
            return r + s * (side * 0.5) * multiplier

        if self.pairs:
            sides = []
            for j, c in enumerate(self.sets):
                idx = pipe.index(f'set{j}/idx', f'set{j}')
                a = pipe.array(f'set{j}/a', c.anchors.subspace, f'set{j}')
                comp = pipe.array(f'set{j}/compliance', c.compliance.subspace, f'set{j}')
                side = pipe.array(f'set{j}/side', S.scalar(), f'set{j}')
                sides.append((idx, a, comp, side))
            for j, (idx, a, comp, side) in enumerate(sides):
                pipe.phase(f'set{j}', apply_side, [s_motor[idx], s_inv[idx], a, comp, side, dt], [s_motor[idx]])
            pipe.phase('body', post, [s_motor, s_old, dt], [s_motor, s_rate])
            for j, (idx, a, comp, side) in enumerate(sides):
                pipe.phase(f'set{j}', v_apply_side, [s_motor[idx], s_rate[idx], s_inv[idx], a, comp, side, dt], [s_rate[idx]])
        per_set = []
        for j, c in enumerate(self.sets if not self.pairs else []):
            i0, i1 = pipe.index(f'set{j}/i0', f'set{j}'), pipe.index(f'set{j}/i1', f'set{j}')
            a0, a1 = pipe.array(f'set{j}/a0', c.anchors.subspace, f'set{j}'), pipe.array(f'set{j}/a1', c.anchors.subspace, f'set{j}')
            map_space = _OperatorKey(c.anchors_map.axes, map_masks[j])
            m0, m1 = pipe.array(f'set{j}/map0', map_space, f'set{j}'), pipe.array(f'set{j}/map1', map_space, f'set{j}')
            comp = pipe.array(f'set{j}/compliance', c.compliance.subspace, f'set{j}')
            per_set.append((i0, i1, a0, a1, m0, m1, comp))
        for j, (i0, i1, a0, a1, m0, m1, comp) in enumerate(per_set):
            pipe.phase(f'set{j}', apply, [s_motor[i0], s_motor[i1], s_inv[i0], s_inv[i1], a0, a1, comp, dt], [s_motor[i0], s_motor[i1]])
        if not self.pairs:
            pipe.phase('body', post, [s_motor, s_old, dt], [s_motor, s_rate])
        for j, (i0, i1, a0, a1, m0, m1, comp) in enumerate(per_set):
            if self.recompute_maps:
                pipe.phase(f'set{j}', v_apply_from_anchors,
                           [s_motor[i0], s_motor[i1], s_rate[i0], s_rate[i1], s_inv[i0], s_inv[i1], a0, a1, comp, dt], [s_rate[i0], s_rate[i1]])
            else:
                pipe.phase(f'set{j}', v_apply, [s_motor[i0], s_motor[i1], s_rate[i0], s_rate[i1], s_inv[i0], s_inv[i1], m0, m1, comp, dt],
                           [s_rate[i0], s_rate[i1]])
        n_loop = 2 + 2 * len(self.sets)
        pipe.phase('body', lambda m, r: (m, r), [s_motor, s_rate], [o_motor, o_rate])
        from numga_b200 import runtime as _rt
        try:
            self.compiled = pipe.build(loop=(1, 1 + n_loop))
        except _rt.NGB200Error as e:
            if 'shared memory' in str(e):
                raise ValueError(f'a block of {block} bodies does not fit one tile ({e}); use FusedStep') from e
            raise
        # per-set operands never change between steps: resolved once
        self._static = {}
        if self.pairs:
            # the reference's [2, c] arrays re-laid out side-major inside each constraint: item 2k + s = side s of constraint k
            from numga_b200.backend import to_soa
            for j, c in enumerate(self.sets):
                nc = int(c.body_idx.shape[1])
                interleave = lambda v: to_soa(v.movedim(0, 1).reshape((2 * nc,) + tuple(v.shape[2:])))
                side = torch.tensor([1.0, -1.0], dtype=ctx.dtype, device=ctx.device).repeat(nc)[:, None]
                self._static.update({
                    f'set{j}/idx': c.body_idx.movedim(0, 1).reshape(-1).contiguous(),
                    f'set{j}/a': ctx.multivector(values=interleave(c.anchors.values), subspace=c.anchors.subspace),
                    f'set{j}/compliance': ctx.multivector(values=to_soa(c.compliance.values.repeat_interleave(2, dim=0)),
                                                          subspace=c.compliance.subspace),
                    f'set{j}/side': ctx.multivector(values=to_soa(side), subspace=S.scalar())})
            self._items = {f'set{j}': 2 * int(c.body_idx.shape[1]) for j, c in enumerate(self.sets)}
            self._items['body'] = n_bodies
            return
        for j, c in enumerate(self.sets):
            self._static.update({f'set{j}/i0': c.body_idx[0].contiguous(), f'set{j}/i1': c.body_idx[1].contiguous(),
                                 f'set{j}/a0': c.anchors[0], f'set{j}/a1': c.anchors[1],
                                 f'set{j}/map0': c.anchors_map[0], f'set{j}/map1': c.anchors_map[1],
                                 f'set{j}/compliance': c.compliance})
        self._items = {f'set{j}': int(c.body_idx.shape[1]) for j, c in enumerate(self.sets)}
        self._items['body'] = n_bodies

    def integrate(self, bodies, dt):
        from numga_b200.backend import soa_empty
        ctx = self.context
        if self.exploit_zeros:
            seen = [(t.data_ptr(), t._version) for t in (bodies.inertia.flat, bodies.inertia_inv.flat)]
            if seen != self._inspected:
                raise ValueError('ChainStep(exploit_zeros=True) was specialised on the zero pattern of other inertia tables; '
                                 'build a new ChainStep for these bodies')
        n = bodies.motor.shape[0]
        motor = ctx.mvtype(ctx, bodies.motor.subspace, soa_empty(len(bodies.motor.subspace), (n,), ctx.dtype, ctx.device))
        rate = ctx.mvtype(ctx, bodies.rate.subspace, soa_empty(len(bodies.rate.subspace), (n,), ctx.dtype, ctx.device))
        arrays = dict(self._static, motor=bodies.motor, rate=bodies.rate, first_moment=bodies.first_moment,
                      inertia=bodies.inertia, inertia_inv=bodies.inertia_inv, damping=bodies.damping, gravity=bodies.gravity,
                      motor_out=motor, rate_out=rate)
        self.compiled.launch(arrays, self._items, repeat=self.unroll, params=[dt])
        return bodies.copy(motor=motor, rate=rate)


_STEPPERS = {}


def step(bodies, constraint_sets, dt, unroll=10):
    """`unroll` sub-steps of `dt / unroll` in ONE launch: the main-loop step of the reference's chain demo
    (numga/examples/physics/run_chain_numpy.py:26-30).  Uses the chain-resident kernel when the system has block
    structure (`ChainStep`), else the six-kernel arrangement (`FusedStep`) sub-step by sub-step.  The stepper is built
    once per (context, constraint sets, unroll) and re-used."""
    ctx = bodies.motor.context
    key = (id(ctx), tuple(id(c) for c in constraint_sets), int(unroll), bodies.motor.shape[0])
    entry = _STEPPERS.get(key)
    if entry is None:
        try:
            stepper = ChainStep(bodies, constraint_sets, unroll=unroll)
        except ValueError:
            stepper = FusedStep(bodies, constraint_sets)
        entry = _STEPPERS[key] = (stepper, ctx, tuple(constraint_sets))       # (keeps the keyed objects alive)
    stepper = entry[0]
    if isinstance(stepper, ChainStep):
        return stepper.integrate(bodies, dt / unroll)
    for _ in range(unroll):
        bodies = stepper.integrate(bodies, dt / unroll)
    return bodies
