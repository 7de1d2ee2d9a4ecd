#!/bin/bash
# cta_group::2 GEMM: tests in their own process with the variant forced on, then bench both ways
mkdir -p gpurun_out
AECF_GEMM_2SM=1 timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 -x > gpurun_out/tests_gemm_2sm.log 2>&1
echo "pytest gemm 2sm exit $?" >> gpurun_out/tests_gemm_2sm.log
AECF_GEMM_2SM=1 timeout 600 python bench.py --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_2sm.json 2> gpurun_out/bench_2sm.err
echo "bench 2sm exit $?" >> gpurun_out/bench_2sm.err
timeout 600 python bench.py --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_1sm.json 2> gpurun_out/bench_1sm.err
tail -30 gpurun_out/tests_gemm_2sm.log; tail -3 gpurun_out/bench_2sm.err
python scripts/show_bench.py gpurun_out/bench_2sm.json 2>/dev/null | head -14; python scripts/show_bench.py gpurun_out/bench_1sm.json 2>/dev/null | head -10
