#!/bin/bash
# callers, smoke(), example step, GEMM ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_callers.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests_callers.log 2>&1
echo "pytest callers exit $?" >> gpurun_out/tests_callers.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python examples/train_vlm.py --steps 20 > gpurun_out/vlm_n1.json 2> gpurun_out/vlm_n1.err
echo "vlm exit $?" >> gpurun_out/vlm_n1.err
timeout 600 python examples/train_vlm.py --steps 20 --heads 8 > gpurun_out/vlm_n1_h8.json 2>> gpurun_out/vlm_n1.err
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 6 -o gpurun_out/prof_gemm \
    python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
tail -25 gpurun_out/tests_callers.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/vlm_n1.json gpurun_out/vlm_n1_h8.json; tail -3 gpurun_out/vlm_n1.err
