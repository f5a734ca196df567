for env in "NGB200_PIPE_UNROLL=0" "NGB200_PIPE_UNROLL=1" "NGB200_PIPE_UNROLL=2" "NGB200_PIPE_UNROLL=1 NGB200_PIPE_PREFETCH=1"; do
 for t in f32 f64; do for u in 1 10; do
  env $env timeout 300 python scripts/profile_chain.py $t 24 $u 2>&1 | tail -1
 done; done; done
