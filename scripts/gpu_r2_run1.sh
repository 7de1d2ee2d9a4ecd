#!/bin/bash
# Round 2, GPU call 1 (ONE box): (1) the default suite -- now including S > 1, bias=False, three-slice rows, the full-size
# oracle test -- (2) the kernel variants built without hardware at the end of round 1, each under its own timeout,
# (3) a same-box A/B of all of them with per-kernel times, (4) the bench line with the new reference baselines.
#   gpurun --timeout 2400 -- bash scripts/gpu_r2_run1.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/r2_1_tests_default.log 2>&1
echo "pytest default exit $?" >> gpurun_out/r2_1_tests_default.log
for v in "2sm_unfixed AECF_GEMM_2SM_FIX=0" "epi2 AECF_GEMM_EPI=2" "epi3 AECF_GEMM_EPI=3" "2sm_ew8 AECF_GEMM_2SM_EW=8" "bwd_stream AECF_POOL_BWD_STREAM=1" "2sm_aux AECF_GEMM_2SM_AUX=1" "apanel AECF_GEMM_APANEL=1" "apanel_ew8 AECF_GEMM_APANEL=1+AECF_GEMM_2SM_EW=8"; do
  set -- $v
  env ${2//+/ } timeout 300 python -m pytest tests/test_gpu_gemm_tcgen05.py tests/test_gpu_parity.py -m gpu -q --tb=short \
      -p no:cacheprovider --timeout 120 -k "gemm or side_output or (bf16 and not full_size) or folded or headline_shape or sharding or b4096" > gpurun_out/r2_1_tests_$1.log 2>&1
  echo "pytest $1 exit $?" >> gpurun_out/r2_1_tests_$1.log
done
for tag in default 2sm_unfixed epi2 epi3 2sm_ew8 epi3_2sm_ew8 all2sm all2sm_ew8 2sm_aux 2sm_aux_ew8 apanel apanel_ew8 apanel_ew8_epi3 bwd_stream default_again; do
  case $tag in default|default_again) E="AECF_NOOP=1";; 2sm_unfixed) E="AECF_GEMM_2SM_FIX=0";;
               epi2) E="AECF_GEMM_EPI=2";; epi3) E="AECF_GEMM_EPI=3";;
               2sm_ew8) E="AECF_GEMM_2SM_EW=8";; epi3_2sm_ew8) E="AECF_GEMM_EPI=3 AECF_GEMM_2SM_EW=8";;
               all2sm) E="AECF_GEMM_2SM=1";; all2sm_ew8) E="AECF_GEMM_2SM=1 AECF_GEMM_2SM_EW=8";;
               2sm_aux) E="AECF_GEMM_2SM_AUX=1";; 2sm_aux_ew8) E="AECF_GEMM_2SM_AUX=1 AECF_GEMM_2SM_EW=8";;
               apanel) E="AECF_GEMM_APANEL=1";; apanel_ew8) E="AECF_GEMM_APANEL=1 AECF_GEMM_2SM_EW=8";;
               apanel_ew8_epi3) E="AECF_GEMM_APANEL=1 AECF_GEMM_2SM_EW=8 AECF_GEMM_EPI=3";;
               bwd_stream) E="AECF_POOL_BWD_STREAM=1";; esac
  env $E timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_1_ab_$tag.json 2> gpurun_out/r2_1_ab_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_1_ab_$tag.json 2>/dev/null | grep -E "^value|kv_proj|d_x|d_kv_weight|out_proj|d_ctx|d_out_weight|pool_bwd|pool_fwd" | cut -c1-110
done
timeout 600 python bench.py > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_1_bench_reference.json 2> gpurun_out/r2_1_bench_reference.err
echo "reference exit $?"; nproc
python - <<'PY'
import json
for f in ("r2_1_bench", "r2_1_bench_reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "reference_gpu_eager", "cpu_baseline")})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -4 gpurun_out/r2_1_tests_*.log
