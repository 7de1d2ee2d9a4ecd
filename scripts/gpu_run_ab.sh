#!/bin/bash
# A/B on one box: ragged-k MMA trimming on / off / on (box-to-box variance is larger than the effect)
mkdir -p gpurun_out
for tag in on1 off on2; do
  v=1; [ "$tag" = "off" ] && v=0
  AECF_GEMM_TRIM_K=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python scripts/show_bench.py gpurun_out/ab_$tag.json 2>/dev/null | grep -E "^value|kv_proj|d_x|d_kv_weight|out_proj" | cut -c1-110
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,temperature.gpu --format=csv
