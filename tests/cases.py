"""Catalogue of hot-path expressions used by the golden-vector generator, the oracle tests and the
GPU parity tests.  Each case is written once against the multivector API (`*`, `^`, `|`, `&`, `>>`,
`.exp()` ...) and therefore runs unchanged on (a) the reference's numpy backend (make_golden.py),
(b) the oracle (oracle/numga_oracle.py) and (c) the B200 backend.

A case: (name, pqr, [(subspace constructor name, 'batch' | 'single')...], input scale, function)
"""

PGA3 = (3, 0, 1)
CGA3 = (4, 1, 0)
STA = (3, 1, 0)
VGA3 = (3, 0, 0)
PGA2 = (2, 0, 1)

B, S = 'batch', 'single'

CASES = [
    # --- config 1: products of batched motors ---------------------------------------------------------
    ('pga3_gp_motor', PGA3, [('even_grade', B), ('even_grade', B)], 1.0, lambda a, b: a * b),
    ('pga3_reverse_product', PGA3, [('even_grade', B), ('even_grade', B)], 1.0, lambda a, b: a.reverse_product(b)),
    ('pga3_squared', PGA3, [('even_grade', B)], 1.0, lambda a: a.squared()),
    ('pga3_symmetric_reverse', PGA3, [('even_grade', B)], 1.0, lambda a: a.symmetric_reverse_product()),
    ('pga3_reverse', PGA3, [('even_grade', B)], 1.0, lambda a: ~a),
    # --- config 2: sandwiches ---------------------------------------------------------------------------------
    ('pga3_sandwich_point_bcast', PGA3, [('even_grade', S), ('antivector', B)], 1.0, lambda r, p: r >> p),
    ('pga3_sandwich_point', PGA3, [('even_grade', B), ('antivector', B)], 1.0, lambda r, p: r >> p),
    ('pga3_sandwich_vector', PGA3, [('even_grade', B), ('vector', B)], 1.0, lambda r, p: r >> p),
    ('pga3_sandwich_bivector', PGA3, [('even_grade', B), ('bivector', B)], 1.0, lambda r, p: r >> p),
    ('pga3_reverse_sandwich_bivector', PGA3, [('even_grade', B), ('bivector', B)], 1.0, lambda r, p: r << p),
    ('pga3_normalize_then_apply', PGA3, [('even_grade', S), ('antivector', B)], 1.0, lambda r, p: r.normalized() >> p),
    # --- graded products ------------------------------------------------------------------------------------------
    ('pga3_wedge_vv', PGA3, [('vector', B), ('vector', B)], 1.0, lambda a, b: a ^ b),
    ('pga3_inner_vb', PGA3, [('vector', B), ('bivector', B)], 1.0, lambda a, b: a | b),
    ('pga3_regressive_pp', PGA3, [('antivector', B), ('antivector', B)], 1.0, lambda a, b: a & b),
    ('pga3_commutator_bb', PGA3, [('bivector', B), ('bivector', B)], 1.0, lambda a, b: a.commutator(b)),
    ('pga3_dual_b', PGA3, [('bivector', B)], 1.0, lambda a: a.dual()),
    ('pga3_gp_full', PGA3, [('multivector', B), ('multivector', B)], 1.0, lambda a, b: a * b),
    ('pga3_wedge_full', PGA3, [('multivector', B), ('multivector', B)], 1.0, lambda a, b: a ^ b),
    ('pga3_inner_full', PGA3, [('multivector', B), ('multivector', B)], 1.0, lambda a, b: a | b),
    ('pga3_regressive_full', PGA3, [('multivector', B), ('multivector', B)], 1.0, lambda a, b: a & b),
    # --- elementwise glue ---------------------------------------------------------------------------------------------
    ('pga3_add_union', PGA3, [('even_grade', B), ('vector', B)], 1.0, lambda a, b: a + b),
    ('pga3_affine', PGA3, [('even_grade', B), ('bivector', B)], 1.0, lambda a, b: a + b * 2 - 1),
    ('pga3_scalar_div', PGA3, [('bivector', B)], 1.0, lambda b: b / 4),
    # --- config 3: exp / normalized / log ----------------------------------------------------------------------------------
    ('pga3_exp', PGA3, [('bivector', B)], 0.5, lambda b: b.exp()),
    ('pga3_normalized', PGA3, [('even_grade', B)], 1.0, lambda m: m.normalized()),
    ('pga3_exp_log', PGA3, [('bivector', B)], 0.5, lambda b: b.exp().motor_log()),
    ('pga3_roundtrip', PGA3, [('bivector', B)], 0.5, lambda b: b.exp().normalized().motor_log()),
    ('pga3_motor_sqrt', PGA3, [('bivector', B)], 0.5, lambda b: b.exp().motor_square_root()),
    ('pga3_inverse_motor', PGA3, [('even_grade', B)], 1.0, lambda m: m.inverse()),
    ('pga3_inverse_vector', PGA3, [('vector', B)], 1.0, lambda m: m.inverse()),
    ('pga3_normalized_vector', PGA3, [('vector', B)], 1.0, lambda m: m.normalized()),
    ('pga3_norm_motor', PGA3, [('even_grade', B)], 1.0, lambda m: m.norm()),
    ('pga3_study_sqrt', PGA3, [('even_grade', B)], 1.0, lambda m: m.symmetric_reverse_product().square_root()),
    ('pga3_exp_degenerate', PGA3, [('bivector', B)], 1.0, lambda b: b.degenerate().exp()),
    # --- config 4: CGA ----------------------------------------------------------------------------------------------------------
    ('cga3_gp_full', CGA3, [('multivector', B), ('multivector', B)], 1.0, lambda a, b: a * b),
    ('cga3_wedge_full', CGA3, [('multivector', B), ('multivector', B)], 1.0, lambda a, b: a ^ b),
    ('cga3_inner_full', CGA3, [('multivector', B), ('multivector', B)], 1.0, lambda a, b: a | b),
    ('cga3_gp_even', CGA3, [('even_grade', B), ('even_grade', B)], 1.0, lambda a, b: a * b),
    ('cga3_sandwich_vector', CGA3, [('even_grade', B), ('vector', B)], 1.0, lambda r, v: r >> v),
    ('cga3_exp', CGA3, [('bivector', B)], 0.3, lambda b: b.exp()),
    # --- decompositions (extension/decompose.py) ------------------------------------------------------------------------------------
    ('pga3_decompose_invariant_l', PGA3, [('bivector', B)], 0.7, lambda b: b.decompose_invariant()[0]),
    ('pga3_decompose_invariant_r', PGA3, [('bivector', B)], 0.7, lambda b: b.decompose_invariant()[1]),
    ('pga3_motor_translator', PGA3, [('even_grade', B)], 1.0, lambda m: m.motor_translator()),
    ('pga3_motor_rotor', PGA3, [('even_grade', B)], 1.0, lambda m: m.motor_rotor()),
    ('pga3_motor_split_t', PGA3, [('bivector', B), ('antivector', B)], 0.5, lambda b, o: b.exp().motor_split(o)[0]),
    ('pga3_motor_split_r', PGA3, [('bivector', B), ('antivector', B)], 0.5, lambda b, o: b.exp().motor_split(o)[1]),
    ('cga3_decompose_invariant_l', CGA3, [('bivector', B)], 0.5, lambda b: b.decompose_invariant()[0]),
    ('vga3_decompose_invariant_l', VGA3, [('bivector', B)], 0.5, lambda b: b.decompose_invariant()[0]),
    # --- other signatures ------------------------------------------------------------------------------------------------------------
    ('sta_gp_full', STA, [('multivector', B), ('multivector', B)], 1.0, lambda a, b: a * b),
    ('sta_exp_log', STA, [('bivector', B)], 0.3, lambda b: b.exp().motor_log()),
    ('vga3_rotor_roundtrip', VGA3, [('bivector', B)], 0.5, lambda b: b.exp().normalized().motor_log()),
    ('vga3_sandwich', VGA3, [('even_grade', B), ('vector', B)], 1.0, lambda r, v: r >> v),
    ('pga2_exp_log', PGA2, [('bivector', B)], 0.5, lambda b: b.exp().motor_log()),
    ('pga2_normalized', PGA2, [('even_grade', B)], 1.0, lambda m: m.normalized()),
]

CASE_BY_NAME = {c[0]: c for c in CASES}
GOLDEN_BATCH = 96


def case_inputs(case, dtype, batch=GOLDEN_BATCH, seed=None):
    """Seeded numpy inputs [batch, n] (or [n] for 'single') for a case; n is resolved by the caller's
    subspace lengths, so this returns a generator function taking the list of component counts."""
    import zlib
    import numpy as np
    name, pqr, specs, scale, _ = case
    seed = zlib.crc32(name.encode()) if seed is None else seed

    def make(widths):
        rng = np.random.default_rng(seed)
        out = []
        for (sub, kind), w in zip(specs, widths):
            shape = (batch, w) if kind == B else (w,)
            out.append((scale * rng.standard_normal(shape)).astype(dtype))
        return out
    return make
