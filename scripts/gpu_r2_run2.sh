#!/bin/bash
# Round 2, GPU call 2: where the projection kernels' time goes.  Each product stand-alone (scripts/gemm_bench.py), every
# kernel family with its main loop or its epilogue removed (AECF_GEMM_DEBUG_SKIP), and cuBLAS on the same shapes.
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 150 python scripts/gemm_bench.py --tag $tag $EXTRA > gpurun_out/r2_2_$tag.json 2> gpurun_out/r2_2_$tag.err || echo "$tag FAILED"; }
EXTRA=--cublas run default AECF_NOOP=1
for sk in 1 2; do run default_skip$sk AECF_GEMM_DEBUG_SKIP=$sk; done
for sk in 0 1 2; do run all1sm_skip$sk AECF_GEMM_2SM=0 AECF_GEMM_DEBUG_SKIP=$sk; done
for sk in 0 2; do run all1sm_nocluster_skip$sk AECF_GEMM_2SM=0 AECF_GEMM_CLUSTER=1 AECF_GEMM_DEBUG_SKIP=$sk; done
for sk in 0 1 2; do run all2sm_skip$sk AECF_GEMM_2SM=1 AECF_GEMM_DEBUG_SKIP=$sk; done
for sk in 0 1 2; do run all2sm_ew8_skip$sk AECF_GEMM_2SM=1 AECF_GEMM_2SM_EW=8 AECF_GEMM_DEBUG_SKIP=$sk; done
for sk in 0 1 2; do run apanel_skip$sk AECF_GEMM_APANEL=1 AECF_GEMM_DEBUG_SKIP=$sk; done
for sk in 0 1 2; do run aux2sm_skip$sk AECF_GEMM_2SM_AUX=1 AECF_GEMM_DEBUG_SKIP=$sk; done
for sk in 0 1; do run epi2_skip$sk AECF_GEMM_EPI=2 AECF_GEMM_DEBUG_SKIP=$sk; done
for sk in 0 1; do run epi3_skip$sk AECF_GEMM_EPI=3 AECF_GEMM_DEBUG_SKIP=$sk; done
python - <<'PY'
import glob, json
names = ("kv_proj", "out_proj", "d_ctx", "d_out_weight", "d_x", "d_kv_weight")
print(f"{'tag':28s}" + "".join(f"{n:>14s}" for n in names))
for f in sorted(glob.glob("gpurun_out/r2_2_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception:
        print(f, "unreadable"); continue
    p = d["products"]
    print(f"{d['tag']:28s}" + "".join(f"{p[n]['us']:14.1f}" if n in p else f"{'-':>14s}" for n in names))
    if d["tag"] == "default":
        print(f"{'  cublas':28s}" + "".join(f"{p[n].get('cublas_us', 0):14.1f}" for n in names))
        print(f"{'  rel_err':28s}" + "".join(f"{p[n].get('rel_err', 0):14.2e}" for n in names))
        print("  kernels:", {n: p[n]["kernel"] for n in names})
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "unsupported_shapes" -p no:cacheprovider 2>&1 | tail -2
