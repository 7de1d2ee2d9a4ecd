#!/bin/bash
# quick confirmation: all GPU tests in one process, smoke, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json 2>/dev/null
