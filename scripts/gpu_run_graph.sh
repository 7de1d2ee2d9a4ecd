#!/bin/bash
# CUDA-graph step: new tests, bench with and without the graph, N=2 data-parallel with NCCL inside the capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graphs.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests_graph.log 2>&1
echo "pytest graph exit $?" >> gpurun_out/tests_graph.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --graph off --no-cpu-baseline --no-e2e > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err
echo "bench exit $?" >> gpurun_out/bench_nograph.err
tail -15 gpurun_out/tests_graph.log; tail -5 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json 2>/dev/null | head -8; python scripts/show_bench.py gpurun_out/bench_nograph.json 2>/dev/null | head -3
