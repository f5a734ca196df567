# tile shapes that fit THREE CTAs per SM (<= 74.6 KB of shared memory, <= 136 registers at 160 threads)
for cfg in "300 160" "288 160" "276 160"; do
 for u in 1 10; do
  NGB200_PIPE_MINB=3 timeout 300 python scripts/profile_chain.py f32 24 $u $cfg 2>&1 | tail -1
 done; done
for u in 1 10; do timeout 300 python scripts/profile_chain.py f32 24 $u 2>&1 | tail -1; done
