#!/bin/bash
# Round 2, GPU call 20: what would the pool backward take without its batch sums, persistent and one sample per warp per CTA?
# (measurement switch AECF_POOL_BWD_NOSUMS=1: the bias gradients are wrong)
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_20_ab_$tag.json 2> gpurun_out/r2_20_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_20_ab_$tag.json 2>/dev/null | grep -E "^value|^roofline  " | cut -c1-90; }
run persistent AECF_POOL_BWD_CHUNK=0
run persistent_nosums AECF_POOL_BWD_CHUNK=0 AECF_POOL_BWD_NOSUMS=1
run chunk8_nosums AECF_POOL_BWD_CHUNK=8 AECF_POOL_BWD_NOSUMS=1
run chunk16_nosums AECF_POOL_BWD_CHUNK=16 AECF_POOL_BWD_NOSUMS=1
run chunk32_nosums AECF_POOL_BWD_CHUNK=32 AECF_POOL_BWD_NOSUMS=1
run chunk32 AECF_POOL_BWD_CHUNK=32
