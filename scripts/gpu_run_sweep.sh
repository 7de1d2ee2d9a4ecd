#!/bin/bash
# check (all GPU tests, smoke, default bench) + BASELINE.json configs[4] stress shapes on one GPU
mkdir -p gpurun_out
bash scripts/gpu_run_check.sh > gpurun_out/check_stdout.txt 2>&1
: > gpurun_out/sweep.jsonl
for shape in "1048576 3 512 8" "262144 8 256 4" "65536 2 2048 16" "131072 4 1024 8" "65536 3 512 1"; do
  set -- $shape
  timeout 200 python bench.py --batch $1 --tokens $2 --dim $3 --heads $4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline \
      >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
  echo "shape $shape exit $?" >> gpurun_out/sweep.err
done
tail -12 gpurun_out/check_stdout.txt
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    if not l.startswith("{"): continue
    d = json.loads(l); c = d["config"]; r = d["roofline"]; f = d["roofline_pool_fwd"]
    print(f"B={c['global_batch']} M={c['tokens']} D={c['embed_dim']} H={c['heads']}: {d['value']/1e6:.1f} M samples/s, {d['ms_per_step']:.3f} ms, "
          f"pool fwd {f['ms']*1e3:.0f} us ({f['frac']:.2f}), bwd {r['ms']*1e3:.0f} us ({r['frac']:.2f})")
PY
grep -v "^$" gpurun_out/sweep.err | tail -8
