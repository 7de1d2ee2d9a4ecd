#!/bin/bash
# First GPU call of round 2, ONE box: (1) the measured default suite, (2) everything built without hardware at the end
# of round 1 -- several queries per sample (csrc/pool_multi.cuh), the pipelined (AECF_GEMM_EPI=2) and eight-warp
# (AECF_GEMM_EPI=3, AECF_GEMM_2SM_EW=8) GEMM epilogues -- each under its own timeout so a hang costs minutes, not the box, and (3) a
# same-box A/B of the three epilogues with per-kernel times.
#   gpurun --timeout 2700 -- bash scripts/gpu_run_round2_first.sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/r2_tests_default.log 2>&1
echo "pytest default exit $?" >> gpurun_out/r2_tests_default.log
AECF_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_multi_query.py tests/test_gpu_parity.py -k "multi_query or three_slices or without_biases" \
    -m gpu -q --tb=short -p no:cacheprovider --timeout 120 > gpurun_out/r2_tests_multi_query.log 2>&1
echo "pytest multi-query exit $?" >> gpurun_out/r2_tests_multi_query.log
for v in "2sm_fix AECF_GEMM_2SM_FIX=1" "epi2 AECF_GEMM_EPI=2" "epi3 AECF_GEMM_EPI=3" "2sm_ew8 AECF_GEMM_2SM_EW=8" "bwd_stream AECF_POOL_BWD_STREAM=1" "2sm_aux AECF_GEMM_2SM_AUX=1" "apanel AECF_GEMM_APANEL=1"; do
  set -- $v
  env $2 timeout 400 python -m pytest tests/test_gpu_gemm_tcgen05.py tests/test_gpu_parity.py -m gpu -q --tb=short \
      -p no:cacheprovider --timeout 120 -k "gemm or side_output or bf16 or folded or headline or sharding" > gpurun_out/r2_tests_$1.log 2>&1
  echo "pytest $1 exit $?" >> gpurun_out/r2_tests_$1.log
done
for tag in epi1 2sm_fix epi2 epi3 2sm_ew8 epi3_2sm_ew8 all2sm_ew8 2sm_aux 2sm_aux_ew8 apanel apanel_ew8 bwd_stream epi1_again; do
  case $tag in epi1|epi1_again) E="AECF_GEMM_EPI=1";; 2sm_fix) E="AECF_GEMM_2SM_FIX=1";;   # the measured CTA-pair kernel with the bulk-group fix epi2) E="AECF_GEMM_EPI=2";; epi3) E="AECF_GEMM_EPI=3";;
               2sm_ew8) E="AECF_GEMM_2SM_EW=8";; epi3_2sm_ew8) E="AECF_GEMM_EPI=3 AECF_GEMM_2SM_EW=8";;
               all2sm_ew8) E="AECF_GEMM_2SM=1 AECF_GEMM_2SM_EW=8";;              # cta_group::2 for every 256-wide product
               2sm_aux) E="AECF_GEMM_2SM_AUX=1";; 2sm_aux_ew8) E="AECF_GEMM_2SM_AUX=1 AECF_GEMM_2SM_EW=8";;
               apanel) E="AECF_GEMM_APANEL=1";; apanel_ew8) E="AECF_GEMM_APANEL=1 AECF_GEMM_2SM_EW=8";;   # resident A panel, K <= 512
               bwd_stream) E="AECF_POOL_BWD_STREAM=1";; esac
  env $E timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_ab_$tag.json 2> gpurun_out/r2_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_ab_$tag.json 2>/dev/null | grep -E "^value|kv_proj|d_x|d_kv_weight|out_proj|d_ctx|d_out_weight|pool_bwd|pool_fwd" | cut -c1-100
done
tail -3 gpurun_out/r2_tests_default.log gpurun_out/r2_tests_multi_query.log gpurun_out/r2_tests_2sm_fix.log gpurun_out/r2_tests_epi2.log gpurun_out/r2_tests_epi3.log gpurun_out/r2_tests_2sm_ew8.log gpurun_out/r2_tests_bwd_stream.log gpurun_out/r2_tests_2sm_aux.log gpurun_out/r2_tests_apanel.log
