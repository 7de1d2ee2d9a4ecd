#!/bin/bash
# Round 2, GPU call 18 (ONE box): the pool kernels on the chunked schedule (CTAs own consecutive samples, handed out by the
# block scheduler) against the persistent one.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2_18_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_18_tests.log; tail -3 gpurun_out/r2_18_tests.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_18_ab_$tag.json 2> gpurun_out/r2_18_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_18_ab_$tag.json 2>/dev/null | grep -E "^value|^roofline" | cut -c1-90; }
run default AECF_NOOP=1
run bwd0 AECF_POOL_BWD_CHUNK=0
run bwd16 AECF_POOL_BWD_CHUNK=16
run bwd64 AECF_POOL_BWD_CHUNK=64
run bwd128 AECF_POOL_BWD_CHUNK=128
run fwd_r1_w8 AECF_POOL_FWD_ROWS=1 AECF_POOL_FWD_WARPS=8
run fwd_r2_w8 AECF_POOL_FWD_ROWS=2 AECF_POOL_FWD_WARPS=8
run fwd_r4_w8 AECF_POOL_FWD_ROWS=4 AECF_POOL_FWD_WARPS=8
run fwd_r8_w8 AECF_POOL_FWD_ROWS=8 AECF_POOL_FWD_WARPS=8
run fwd_r4_w24 AECF_POOL_FWD_ROWS=4 AECF_POOL_FWD_WARPS=24
run fwd_r4_w4 AECF_POOL_FWD_ROWS=4 AECF_POOL_FWD_WARPS=4
run default_again AECF_NOOP=1
