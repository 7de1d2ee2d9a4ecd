#!/bin/bash
# GEMM epilogue experiment: tests (default heuristic and 2SM forced off), then bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 -x > gpurun_out/tests_gemm.log 2>&1
echo "pytest gemm exit $?" >> gpurun_out/tests_gemm.log
AECF_GEMM_2SM=0 timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 -x > gpurun_out/tests_gemm_1sm.log 2>&1
echo "pytest gemm 1sm exit $?" >> gpurun_out/tests_gemm_1sm.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callers.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
timeout 600 python bench.py --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err
AECF_GEMM_2SM=0 timeout 600 python bench.py --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_1sm.json 2> gpurun_out/bench_1sm.err
tail -3 gpurun_out/tests_gemm.log; tail -3 gpurun_out/tests_gemm_1sm.log; tail -3 gpurun_out/tests.log
python scripts/show_bench.py gpurun_out/bench_auto.json 2>/dev/null | head -16; python scripts/show_bench.py gpurun_out/bench_1sm.json 2>/dev/null | head -14
