#!/bin/bash
# as the driver does it (all GPU tests in one process, smoke, default bench, reference arm), then the ncu evidence:
# launch list of the eager step and `--set full` captures of the pool kernels and of the GEMMs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
NCU_ARGS="bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --graph off"
timeout 200 python $NCU_ARGS > gpurun_out/plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python $NCU_ARGS > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pool_(fwd|bwd)" -s 6 -c 3 -o gpurun_out/prof_pool_fold \
    python $NCU_ARGS > gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 12 -c 6 -o gpurun_out/prof_gemm_fold \
    python $NCU_ARGS > gpurun_out/ncu_gemm.log 2>&1
tail -4 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json 2>/dev/null; head -c 300 gpurun_out/bench_reference.json; ls -la gpurun_out/*.ncu-rep
