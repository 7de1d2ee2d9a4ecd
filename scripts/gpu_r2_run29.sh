#!/bin/bash
# Round 2, GPU call 29 (ONE box): streaming forward with unrolled head butterflies and the dense specialisation.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2_29_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_29_tests.log; tail -3 gpurun_out/r2_29_tests.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_29_ab_$tag.json 2> gpurun_out/r2_29_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_29_ab_$tag.json 2>/dev/null | grep -E "^value|^roofline" | cut -c1-90; }
run dense AECF_NOOP=1
run generic AECF_POOL_FWD_DENSE=0
run dense_again AECF_NOOP=1
