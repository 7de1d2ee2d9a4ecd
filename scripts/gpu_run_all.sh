#!/bin/bash
# everything as the driver does it: all gpu tests in one process, smoke, default bench; plus PDL off for comparison
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
AECF_PDL=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -4 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json 2>/dev/null; python scripts/show_bench.py gpurun_out/bench_nopdl.json 2>/dev/null | head -3; head -c 400 gpurun_out/bench_reference.json
