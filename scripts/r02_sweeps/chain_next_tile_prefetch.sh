# next-tile bulk L2 prefetch (NGB200_PIPE_PREFETCH=2) against the default
for pf in 2 0; do for t in f32 f64; do for u in 1 10; do
  NGB200_PIPE_PREFETCH=$pf timeout 300 python scripts/profile_chain.py $t 24 $u 2>&1 | tail -1
done; done; done
