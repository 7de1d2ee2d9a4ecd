#!/bin/bash
# A/B on one box: default GEMM kernel choice vs cta_group::2 forced for every 256-wide product vs no cluster multicast
mkdir -p gpurun_out
for tag in default 2sm nocluster; do
  case $tag in default) E="";; 2sm) E="AECF_GEMM_2SM=1";; nocluster) E="AECF_GEMM_CLUSTER=1";; esac
  env $E timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ab2_$tag.json 2> gpurun_out/ab2_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/ab2_$tag.json 2>/dev/null | grep -E "^value|kv_proj|d_x|d_kv_weight|out_proj|d_ctx|d_out_weight" | cut -c1-100
done
