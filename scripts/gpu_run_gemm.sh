#!/bin/bash
# GEMM-focused run: tcgen05 tests alone (a trap poisons the context), then all tests, then bench with and without clusters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests_gemm.log 2>&1
echo "pytest gemm exit $?" >> gpurun_out/tests_gemm.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
timeout 600 python bench.py --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_cluster2.json 2> gpurun_out/bench_cluster2.err
AECF_GEMM_CLUSTER=1 timeout 600 python bench.py --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_cluster1.json 2> gpurun_out/bench_cluster1.err
tail -4 gpurun_out/tests_gemm.log; tail -4 gpurun_out/tests.log
python scripts/show_bench.py gpurun_out/bench_cluster2.json; python scripts/show_bench.py gpurun_out/bench_cluster1.json | head -14
