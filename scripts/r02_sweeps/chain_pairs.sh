# ChainStep(pairs=True): two threads per constraint (one per side, warp-shuffle exchange) against the default
export CHAIN_PAIRS=1
for u in 1 10; do
  for cfg in "f32 1 384" "f32 2 384" "f32 2 240" "f64 1 192" "f64 2 192" "f64 3 96"; do
    set -- $cfg
    NGB200_PIPE_MINB=$2 timeout 300 python scripts/profile_chain.py $1 24 $u $3 2>&1 | tail -1 | cut -c1-420
  done
done
