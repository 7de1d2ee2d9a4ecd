#!/bin/bash
# Round 2, GPU call 23 (ONE box): streaming pool forward with a 3-row cp.async ring per warp against the 2-row ring.
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_23_ab_$tag.json 2> gpurun_out/r2_23_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_23_ab_$tag.json 2>/dev/null | grep -E "^value|^roofline" | cut -c1-90; }
run stages2 AECF_POOL_FWD_STAGES=2
run stages3 AECF_POOL_FWD_STAGES=3
run stages2_again AECF_POOL_FWD_STAGES=2
run stages3_again AECF_POOL_FWD_STAGES=3
AECF_POOL_FWD_STAGES=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2_23_tests_stages3.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_23_tests_stages3.log; tail -3 gpurun_out/r2_23_tests_stages3.log
