#!/bin/bash
# GPU run 2: tcgen05 GEMM tests in their own process (a trap poisons the context), then everything else
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests_gemm.log 2>&1
echo "pytest gemm exit $?" >> gpurun_out/tests_gemm.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -15 gpurun_out/tests_gemm.log; tail -8 gpurun_out/tests.log; tail -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
