#!/bin/bash
# GPU run 3: all tests in one process (as the driver does), bench, then ncu launch list + pool capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pool_(fwd|bwd)" -s 3 -c 3 -o gpurun_out/prof_pool \
    python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -6 gpurun_out/tests.log; tail -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
