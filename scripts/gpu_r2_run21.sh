#!/bin/bash
# Round 2, GPU call 21 (ONE box): the pool backward without batch sums (per-sample [s | sum ds] -> the tail forms the bias
# sums next to the [dWv;R] product) on the chunked schedule; parity, then the A/B of samples per CTA.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2_21_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_21_tests.log; tail -3 gpurun_out/r2_21_tests.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_21_ab_$tag.json 2> gpurun_out/r2_21_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_21_ab_$tag.json 2>/dev/null | grep -E "^value|^roofline  |d_kv_weight|d_x|grad_gather" | cut -c1-90; }
run default AECF_NOOP=1
run persistent AECF_POOL_BWD_CHUNK=0
run chunk8 AECF_POOL_BWD_CHUNK=8
run chunk16 AECF_POOL_BWD_CHUNK=16
run chunk64 AECF_POOL_BWD_CHUNK=64
run noside AECF_SIDE_STREAM=0
run default_again AECF_NOOP=1
