#!/bin/bash
# Round 2, GPU call 4 (ONE box): the gradient tail as four parallel kernels; A/B side stream and fused loss.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2_4_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_4_tests.log
tail -4 gpurun_out/r2_4_tests.log
for tag in default noside noloss noside_noloss default_again; do
  case $tag in noside) E="AECF_SIDE_STREAM=0";; noloss) E="AECF_FUSED_LOSS=0";; noside_noloss) E="AECF_SIDE_STREAM=0 AECF_FUSED_LOSS=0";; *) E="AECF_NOOP=1";; esac
  env $E timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_4_ab_$tag.json 2> gpurun_out/r2_4_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_4_ab_$tag.json 2>/dev/null | grep -v "^pool only\|^roofline  " | cut -c1-100
done
