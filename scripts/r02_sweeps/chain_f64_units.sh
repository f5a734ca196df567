# fp64 record layout in 8-byte units (odd count per record + one unit per 16 rows) against 16-byte chunks
timeout 600 python -m pytest tests/test_physics.py -m gpu -x -q 2>&1 | tail -2
for units in 1 0; do for u in 1 10; do
  NGB200_PIPE_F64_UNITS=$units timeout 300 python scripts/profile_chain.py f64 24 $u 2>&1 | tail -1 | cut -c1-330
done; done
