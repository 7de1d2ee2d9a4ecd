#!/bin/bash
# Round 2, GPU call 22 (ONE box): the pool backward with no batch sums at all (no dropout: d_bias_v = Wo^T colsum(d_out),
# d_bias_k = 0, formed by the tail) on the chunked schedule; parity, then the A/B of samples per CTA.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2_22_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_22_tests.log; tail -3 gpurun_out/r2_22_tests.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_22_ab_$tag.json 2> gpurun_out/r2_22_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_22_ab_$tag.json 2>/dev/null | grep -E "^value|^roofline  |d_kv_weight|d_x|grad_gather" | cut -c1-90; }
run default AECF_NOOP=1
run persistent AECF_POOL_BWD_CHUNK=0
run chunk8 AECF_POOL_BWD_CHUNK=8
run chunk16 AECF_POOL_BWD_CHUNK=16
run chunk64 AECF_POOL_BWD_CHUNK=64
run noside AECF_SIDE_STREAM=0
run default_again AECF_NOOP=1
