#!/bin/bash
# Round 2, GPU call 13 (ONE box): aligned loss hand-over in the forward, two-stage fork of the gradient tail, batched loads
# in the small kernels; A/B and an ncu look at the small kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2_13_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_13_tests.log
tail -4 gpurun_out/r2_13_tests.log
for tag in default noside noloss default_again; do
  case $tag in noside) E="AECF_SIDE_STREAM=0";; noloss) E="AECF_FUSED_LOSS=0";; *) E="AECF_NOOP=1";; esac
  env $E timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_13_ab_$tag.json 2> gpurun_out/r2_13_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_13_ab_$tag.json 2>/dev/null | grep -v "^pool only\|^roofline  " | cut -c1-100
done
AECF_SIDE_STREAM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grad_|fold_prepare|pool_fwd_stream" --launch-skip 6 -c 6 -f -o gpurun_out/r2_13_small \
    python bench.py --steps 2 --warmup 1 --graph off --no-e2e --no-cpu-baseline > gpurun_out/r2_13_ncu_small.log 2>&1
ncu -i gpurun_out/r2_13_small.ncu-rep --page raw --csv > gpurun_out/r2_13_small_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_13_small_raw.csv > gpurun_out/r2_13_small_summary.txt 2>&1
grep -E "=====|gpu__time_duration|dram__bytes|grid_size|issue_active|stalls" gpurun_out/r2_13_small_summary.txt | cut -c1-170
