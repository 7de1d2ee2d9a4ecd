#!/bin/bash
# Round 2, GPU call 25 (ONE box): panel mode of the single-CTA tcgen05 GEMM (a row block's A k-blocks loaded once for all
# of its N tiles).  GEMM tests (with the forced-panel child) first, then the stand-alone products, then the step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -p no:cacheprovider --timeout 700 -x > gpurun_out/r2_25_tests_gemm.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_25_tests_gemm.log; tail -3 gpurun_out/r2_25_tests_gemm.log
grep -q "pytest exit 0" gpurun_out/r2_25_tests_gemm.log || exit 1
: > gpurun_out/r2_25_gemm_panel.jsonl
AECF_GEMM_PANEL=0 timeout 120 python scripts/gemm_bench.py --tag "panel_off" >> gpurun_out/r2_25_gemm_panel.jsonl 2> gpurun_out/r2_25_gemm_err.txt || echo "failed off"
timeout 120 python scripts/gemm_bench.py --cublas --tag "panel_on" >> gpurun_out/r2_25_gemm_panel.jsonl 2>> gpurun_out/r2_25_gemm_err.txt || echo "failed on"
AECF_GEMM_STAGES=8,2 timeout 120 python scripts/gemm_bench.py --tag "panel_on_b2" >> gpurun_out/r2_25_gemm_panel.jsonl 2>> gpurun_out/r2_25_gemm_err.txt || echo "failed b2"
AECF_GEMM_DEBUG_SKIP=2 timeout 120 python scripts/gemm_bench.py --tag "panel_on_mainloop_only" >> gpurun_out/r2_25_gemm_panel.jsonl 2>> gpurun_out/r2_25_gemm_err.txt || echo "failed skip2"
AECF_GEMM_DEBUG_SKIP=1 timeout 120 python scripts/gemm_bench.py --tag "panel_on_epilogue_only" >> gpurun_out/r2_25_gemm_panel.jsonl 2>> gpurun_out/r2_25_gemm_err.txt || echo "failed skip1"
python - <<'P'
import json
for l in open('gpurun_out/r2_25_gemm_panel.jsonl'):
    try: d = json.loads(l)
    except Exception: continue
    res = d.get('products', d)
    print(d.get('tag'), {k: (round(v['us'], 1), round(v.get('cublas_us', 0), 1), v.get('kernel', '')[-14:]) for k, v in res.items() if isinstance(v, dict) and 'us' in v})
P
for tag in panel_on panel_off; do
  case $tag in panel_off) E="AECF_GEMM_PANEL=0";; *) E="AECF_NOOP=1";; esac
  env $E timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_25_bench_$tag.json 2> gpurun_out/r2_25_bench_$tag.err
  echo "== $tag"; python scripts/show_bench.py gpurun_out/r2_25_bench_$tag.json 2>/dev/null | grep -v "^pool only\|^host" | cut -c1-100
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x -k "full_size or bf16" > gpurun_out/r2_25_tests_parity.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_25_tests_parity.log; tail -3 gpurun_out/r2_25_tests_parity.log
