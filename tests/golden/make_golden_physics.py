"""Golden data for the rigid-body physics core step (BASELINE.json config 5), from the UNMODIFIED reference:
`setup_bodies(NumpyContext('x+y+z+w0', float64), n_bodies=12)` then `bodies.integrate(dt, constraint_sets)`
(numga/examples/physics/{setup_chain,core,base}.py).  Writes tests/golden/physics.npz:
  initial state, the state after pre_integrate / each constraint set / post_integrate / each velocity set of the
  first sub-step, and the state after 1, 2 and 5 full sub-steps.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

load_reference()
from numga.backend.numpy.context import NumpyContext  # noqa: E402
from numga.examples.physics.setup_chain import setup_bodies  # noqa: E402


def state(prefix, bodies, out):
    out[f'{prefix}/motor'] = np.array(bodies.motor.values)
    out[f'{prefix}/rate'] = np.array(bodies.rate.values)


def main():
    out = {}
    for tag, dtype in (('f64', np.float64), ('f32', np.float32)):
        ctx = NumpyContext('x+y+z+w0', dtype=dtype)
        bodies, sets = setup_bodies(ctx, n_bodies=12)
        dt = 1 / 200
        pre = f'{tag}'
        out[f'{pre}/init/motor'] = np.array(bodies.motor.values)
        out[f'{pre}/init/rate'] = np.array(bodies.rate.values)
        out[f'{pre}/init/first_moment'] = np.array(bodies.first_moment.values)
        out[f'{pre}/init/inertia'] = np.array(bodies.inertia.kernel)
        out[f'{pre}/init/inertia_inv'] = np.array(bodies.inertia_inv.kernel)
        out[f'{pre}/init/damping'] = np.array(bodies.damping.values)
        out[f'{pre}/init/gravity'] = np.array(bodies.gravity.values)
        out[f'{pre}/init/gravity_blades'] = bodies.gravity.subspace.blades.astype(np.int64)
        for k, c in enumerate(sets):
            out[f'{pre}/set{k}/body_idx'] = np.array(c.body_idx)
            out[f'{pre}/set{k}/anchors'] = np.array(c.anchors.values)
            out[f'{pre}/set{k}/compliance'] = np.array(c.compliance.values)
            out[f'{pre}/set{k}/anchors_map'] = np.array(c.anchors_map.kernel)
        # a perturbed start makes every term of the step non-trivial
        rng = np.random.default_rng(7)
        bodies = bodies.copy(rate=bodies.rate + ctx.multivector.bivector(0.3 * rng.standard_normal((12, 6))))
        out[f'{pre}/start/rate'] = np.array(bodies.rate.values)
        new = bodies.pre_integrate(dt); state(f'{pre}/pre', new, out)
        for k, c in enumerate(sets):
            new = c.apply(new, dt=dt); state(f'{pre}/apply{k}', new, out)
        new = new.post_integrate(bodies, dt=dt); state(f'{pre}/post', new, out)
        for k, c in enumerate(sets):
            new = c.v_apply(new, dt=dt); state(f'{pre}/vapply{k}', new, out)
        b = bodies
        for step in range(1, 6):
            b = b.integrate(dt, sets)
            if step in (1, 2, 5):
                state(f'{pre}/step{step}', b, out)
        assert np.array_equal(out[f'{pre}/step1/motor'], out[f'{pre}/vapply1/motor'])
    tennis(out)
    np.savez_compressed(os.path.join(HERE, 'physics.npz'), **out)
    print('wrote', len(out), 'arrays')



def tennis(out):
    """numga/examples/tennis_racket_theorem.py: free tumbling of point clouds with distinct moments of inertia, one body
    per spin plane, in 3 and 4 euclidean dimensions (no constraints, no degenerate direction)."""
    from numga.algebra.algebra import Algebra
    from numga.examples.physics.core import Body
    for pqr in ((3, 0, 0), (4, 0, 0)):
        ctx = NumpyContext(Algebra.from_pqr(*pqr), dtype=np.float64)
        n = ctx.algebra.description.n_dimensions
        cube = 2 * ((np.arange(2 ** n)[:, None] & (1 << np.arange(n))) > 0) - 1
        points = ctx.multivector.vector(cube * (np.arange(n) + 1)).dual()
        nb = len(ctx.subspace.bivector())
        body = Body.from_point_cloud(points=points.repeat('... -> b ...', b=nb))
        jitter = np.random.default_rng(3).standard_normal((nb, nb)) * 1e-5
        body.rate = body.rate + ctx.multivector.bivector(np.eye(nb) + jitter)
        tag = 'tennis%d' % n
        out[f'{tag}/jitter'] = jitter
        out[f'{tag}/energy0'] = np.array(body.kinetic_energy().values)
        for step in range(1, 9):
            body = body.integrate(1 / 4)
            if step in (1, 8):
                out[f'{tag}/step{step}/motor'] = np.array(body.motor.values)
                out[f'{tag}/step{step}/rate'] = np.array(body.rate.values)
        out[f'{tag}/energy8'] = np.array(body.kinetic_energy().values)


if __name__ == '__main__':
    main()
