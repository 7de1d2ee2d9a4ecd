#!/bin/bash
# Round 2, GPU call 30 (ONE box): the evidence run -- full GPU test suite, smoke(), the full default bench line (with
# reference_gpu_eager, cpu_baseline, cuBLAS yardstick, e2e), the reference arm, the ncu launch list and one --set full capture.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/r2_30_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_30_tests.log
tail -4 gpurun_out/r2_30_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_30_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2_30_smoke.log
timeout 600 python bench.py > gpurun_out/r2_30_bench.json 2> gpurun_out/r2_30_bench.err; echo "bench exit $?"
python scripts/show_bench.py gpurun_out/r2_30_bench.json 2>/dev/null | head -12 | cut -c1-160
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_30_bench_reference.json 2> gpurun_out/r2_30_bench_reference.err; echo "reference exit $?"
cut -c1-400 gpurun_out/r2_30_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_30_launches.csv \
    python bench.py --steps 2 --warmup 1 --graph off --no-e2e --no-cpu-baseline > gpurun_out/r2_30_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_|pool_|grad_|fold_" --launch-skip 14 -c 16 -f -o gpurun_out/r2_30_prof \
    python bench.py --steps 2 --warmup 1 --graph off --no-e2e --no-cpu-baseline > gpurun_out/r2_30_ncu_full.log 2>&1
ls -la gpurun_out/r2_30_prof.ncu-rep
ncu -i gpurun_out/r2_30_prof.ncu-rep --page raw --csv > gpurun_out/r2_30_prof_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_30_prof_raw.csv > gpurun_out/r2_30_ncu_summary.txt 2>&1
grep -E "=====|gpu__time_duration|dram__bytes_(read|write).sum " gpurun_out/r2_30_ncu_summary.txt | cut -c1-150
