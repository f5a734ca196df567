"""Edge-case inputs for the non-linear chains (zero bivector, identity motor, pure translator, pure
rotor, |b| near pi, null / tiny / huge motors).  Shared by the golden generator (reference), the oracle test and
the GPU parity test.  Each entry: (name, input subspace, rows, function)."""
import numpy as np

PGA3 = (3, 0, 1)


def _bivectors():
    rng = np.random.default_rng(42)
    rot = rng.standard_normal(3)
    rot /= np.linalg.norm(rot)
    rows = [np.zeros(6)]                                              # zero bivector -> identity motor
    rows += [np.r_[0, 0, 0, t] for t in (np.eye(3) * 2.5)]            # pure translations (degenerate part only: xw,yw,zw)
    rows += [np.r_[a * rot, 0, 0, 0] for a in (1e-8, 1e-3, 0.5, 1.5)]  # pure rotations, tiny to large angle
    rows += [np.r_[a * rot, 0.3, -0.2, 0.1] for a in (np.pi / 2 - 1e-3, np.pi / 2 - 1e-6, 3.0)]   # motor angle near pi
    rows += [np.r_[1e-20 * rot, 1e-20, 0, 0], np.r_[20 * rot, 5, 5, 5]]            # tiny and huge
    return np.array(rows)


def _motors():
    rng = np.random.default_rng(43)
    ident = np.r_[1.0, np.zeros(7)]
    rows = [ident, -ident, 3.0 * ident]                               # identity, its double cover twin, unnormalized
    rows += [np.r_[1, 0, 0, 0, 0.5, -1.5, 2.0, 0]]                    # pure translator
    q = rng.standard_normal(4); q /= np.linalg.norm(q)
    rows += [np.r_[q, 0, 0, 0, 0]]                                    # pure rotor
    rows += [np.r_[0, q[1:] / np.linalg.norm(q[1:]), 0, 0, 0, 0]]     # rotation by pi (scalar part 0)
    rows += [np.zeros(8)]                                             # null: norm 0
    rows += [np.r_[0, 0, 0, 0, 1, 2, 3, 4.0]]                         # degenerate only: norm 0, non-zero components
    rows += [1e-20 * rng.standard_normal(8), 1e15 * rng.standard_normal(8)]      # tiny / huge
    rows += [rng.standard_normal(8) for _ in range(4)]                # generic unnormalized
    return np.array(rows)


EDGE_CASES = [
    ('exp', 'bivector', _bivectors, lambda b: b.exp()),
    ('exp_log', 'bivector', _bivectors, lambda b: b.exp().motor_log()),
    ('roundtrip', 'bivector', _bivectors, lambda b: b.exp().normalized().motor_log()),
    ('normalized', 'even_grade', _motors, lambda m: m.normalized()),
    ('normalized_log', 'even_grade', _motors, lambda m: m.normalized().motor_log()),
    ('motor_sqrt', 'even_grade', _motors, lambda m: m.normalized().motor_square_root()),
    ('inverse', 'even_grade', _motors, lambda m: m.inverse()),
    ('norm', 'even_grade', _motors, lambda m: m.norm()),
]
