#!/bin/bash
# Round 2, GPU call 3 (ONE box): the restructured step -- gradient tail on a side stream, fused entropy_loss, merged
# prepare, pruned GEMMs -- tests, bench, the side-stream A/B, and the ncu captures of the SHIPPED kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/r2_3_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_3_tests.log
tail -5 gpurun_out/r2_3_tests.log
for tag in default noside default_again; do
  case $tag in noside) E="AECF_SIDE_STREAM=0";; *) E="AECF_NOOP=1";; esac
  env $E timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_3_ab_$tag.json 2> gpurun_out/r2_3_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_3_ab_$tag.json 2>/dev/null | cut -c1-120
done
timeout 600 python bench.py > gpurun_out/r2_3_bench.json 2> gpurun_out/r2_3_bench.err; echo "bench exit $?"
python scripts/show_bench.py gpurun_out/r2_3_bench.json 2>/dev/null | head -8
# ncu: launch list of one eager step region, then --set full on the step's kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_3_launches.csv \
    python bench.py --steps 2 --warmup 1 --graph off --no-e2e --no-cpu-baseline > gpurun_out/r2_3_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 13 -c 13 -f -o gpurun_out/r2_3_prof \
    python bench.py --steps 2 --warmup 1 --graph off --no-e2e --no-cpu-baseline > gpurun_out/r2_3_ncu_full.log 2>&1
ls -la gpurun_out/r2_3_prof.ncu-rep
ncu -i gpurun_out/r2_3_prof.ncu-rep --page raw --csv > gpurun_out/r2_3_prof_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_3_prof_raw.csv > gpurun_out/r2_3_ncu_summary.txt 2>&1
grep -E "=====|gpu__time_duration|dram__bytes|tensor_cycles|dram_throughput" gpurun_out/r2_3_ncu_summary.txt | cut -c1-150
