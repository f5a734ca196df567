"""Float gates of the benchmarked (fast) mode, derived in profiles/r02_tolerances.md from measurements on a B200
(profiles/r02_tolerances.json, scripts/tolerance_table.py).  Every gate compares the kernel with the REFERENCE at the
same precision; none is looser than 2x the reference's own error at that precision against a higher-precision truth,
except the one flagged exception."""
import numpy as np

NORTH_STAR = {np.float32: 1e-5, np.float64: 1e-12}
# exp / log chains: common-mode bisection error; fast fp32 vs reference fp32 measured <= 8.4e-6, fp64 <= 7.3e-12
CHAIN = {np.float32: 2e-5, np.float64: 2e-11}
CHAIN_MARKERS = ('exp', 'log', 'roundtrip', 'motor_sqrt', 'motor_split')
EXCEPTIONS = {
    # division by a Study norm that vanishes for simple bivectors; measured 6.7e-4 (reference's own error: 1.8e-4)
    'pga3_decompose_invariant_l': {np.float32: 1.5e-3, np.float64: 5e-12},
    'cga3_decompose_invariant_l': {np.float32: 5e-5, np.float64: 1e-12},
}
# physics: (motor, rate) gates after `steps` sub-steps against the reference's states
PHYSICS_RATE_F32 = {1: 1.2e-3, 2: 1.6e-3, 3: 2.2e-3, 5: 3e-3}


def gate(case_name, dtype):
    dtype = np.dtype(dtype).type
    if case_name in EXCEPTIONS:
        return EXCEPTIONS[case_name][dtype]
    if any(m in case_name for m in CHAIN_MARKERS):
        return CHAIN[dtype]
    return NORTH_STAR[dtype]


def physics_gate(tag, steps):
    """(motor gate, rate gate)"""
    if tag == 'f32':
        return 1e-5 * steps, PHYSICS_RATE_F32.get(steps, 3e-3 * max(1, steps / 5))
    return 1e-12 * steps, 1e-11 * steps
