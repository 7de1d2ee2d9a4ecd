#!/bin/bash
# folded key projection: new tests first (fast feedback), then the whole GPU suite, smoke, bench folded / unfolded
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 \
   -k "folded or side_output or bf16 or wide" > gpurun_out/tests_fold.log 2>&1
echo "pytest fold exit $?" >> gpurun_out/tests_fold.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --fold off --no-cpu-baseline --no-e2e > gpurun_out/bench_unfolded.json 2> gpurun_out/bench_unfolded.err
tail -15 gpurun_out/tests_fold.log; tail -6 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json 2>/dev/null; python scripts/show_bench.py gpurun_out/bench_unfolded.json 2>/dev/null | head -3
