"""numga_b200 — B200 (sm_100a) backend for numga's batched multivector operators.

    from numga_b200 import B200Context
    ctx = B200Context((3, 0, 1))                       # 3D PGA, float32, cuda:0
    R = ctx.multivector.motor(values8)                 # one motor
    p = ctx.multivector.antivector(values_n_by_4)      # a batch of points, stored structure-of-arrays in HBM
    q = R >> p                                         # one generated kernel, motor pre-contracted in registers
"""
from numga_b200.algebra import Algebra  # noqa: F401
from numga_b200.backend import B200Context, B200MultiVector, B200Operator  # noqa: F401
