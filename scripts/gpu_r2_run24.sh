#!/bin/bash
# Round 2, GPU call 24 (ONE box): the single-CTA tcgen05 GEMM with separate A (deep) and B (shallow) operand rings.
# GEMM tests first (every wait traps after ~2 s instead of hanging), then the stand-alone products over ring geometries.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -p no:cacheprovider --timeout 120 -x > gpurun_out/r2_24_tests_gemm.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_24_tests_gemm.log; tail -3 gpurun_out/r2_24_tests_gemm.log
grep -q "pytest exit 0" gpurun_out/r2_24_tests_gemm.log || exit 1
: > gpurun_out/r2_24_gemm_rings.jsonl
for st in 4,4 6,3 7,3 6,4 8,3 8,2 5,4 7,2 9,2 10,2; do
  AECF_GEMM_STAGES=$st timeout 120 python scripts/gemm_bench.py --tag "stages_$st" >> gpurun_out/r2_24_gemm_rings.jsonl 2> gpurun_out/r2_24_gemm_err.txt || echo "failed $st"
done
timeout 120 python scripts/gemm_bench.py --cublas --tag default >> gpurun_out/r2_24_gemm_rings.jsonl 2>> gpurun_out/r2_24_gemm_err.txt
python - <<'P'
import json
for l in open('gpurun_out/r2_24_gemm_rings.jsonl'):
    try: d = json.loads(l)
    except Exception: continue
    res = d.get('results', d)
    row = {k: (round(v['us'], 1), round(v.get('cublas_us', 0), 1), '%.0e' % v.get('rel_err', 0)) for k, v in res.items() if isinstance(v, dict) and 'us' in v}
    print(d.get('tag'), row)
P
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_24_bench.json 2> gpurun_out/r2_24_bench.err
python scripts/show_bench.py gpurun_out/r2_24_bench.json 2>/dev/null | cut -c1-100
