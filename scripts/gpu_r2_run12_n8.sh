#!/bin/bash
# Round 2, GPU call 12 (EIGHT GPUs, one box): the in-backward gradient sum at 8 ranks, weak and strong scaling at
# N = 1, 2, 4, 8, the config-5 sweep with parity at 8 x 131 072 rows, the VLM step at 1/2/4/8.
mkdir -p gpurun_out
P=29600
tr() { n=$1; shift; P=$((P+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; }
nvidia-smi topo -m > gpurun_out/r2_12_topo.txt 2>&1
tr 8 scripts/dp_check.py > gpurun_out/r2_12_dp_check_n8.log 2>&1; echo "dp_check n8 exit $?"
grep -E "^\{" gpurun_out/r2_12_dp_check_n8.log | cut -c1-400
tr 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --dp-study > gpurun_out/r2_12_weak_n8_study.json 2> gpurun_out/r2_12_weak_n8_study.err; echo "weak n8 study exit $?"
tr 8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_12_weak_n8.json 2> gpurun_out/r2_12_weak_n8.err; echo "weak n8 exit $?"
for n in 4 2; do
  tr $n bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_12_weak_n$n.json 2> gpurun_out/r2_12_weak_n$n.err; echo "weak n$n exit $?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_12_weak_n1.json 2> gpurun_out/r2_12_weak_n1.err; echo "weak n1 exit $?"
for n in 8 4 2; do
  tr $n bench.py --gpus $n --global-batch 65536 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_12_strong_n$n.json 2> gpurun_out/r2_12_strong_n$n.err; echo "strong n$n exit $?"
done
tr 8 tests/sweep_parity.py > gpurun_out/r2_12_sweep_n8.jsonl 2> gpurun_out/r2_12_sweep_n8.err; echo "sweep n8 exit $?"
grep -v "^{" gpurun_out/r2_12_sweep_n8.jsonl
for n in 8 4 2; do
  tr $n examples/train_vlm.py --steps 20 > gpurun_out/r2_12_vlm_n$n.json 2> gpurun_out/r2_12_vlm_n$n.err; echo "vlm n$n exit $?"
done
timeout 200 python examples/train_vlm.py --steps 20 > gpurun_out/r2_12_vlm_n1.json 2> gpurun_out/r2_12_vlm_n1.err; echo "vlm n1 exit $?"
python - <<'PY'
import json
def last(f):
    try:
        return json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    except Exception as e:
        return None
base = last("r2_12_weak_n1")
for n in (1, 2, 4, 8):
    d = last(f"r2_12_weak_n{n}")
    if d: print(f"weak   N={n}: {d['value']/1e6:8.1f} M samples/s  {d['ms_per_step']:.4f} ms  x{d['value']/base['value']:.2f}", (d.get('data_parallel') or {}).get('rank_spread'))
for n in (2, 4, 8):
    d = last(f"r2_12_strong_n{n}")
    if d: print(f"strong N={n}: {d['value']/1e6:8.1f} M samples/s  {d['ms_per_step']:.4f} ms  x{d['value']/base['value']:.2f}")
d = last("r2_12_weak_n8_study")
if d:
    dp = d["data_parallel"]
    print("study N=8:", d["ms_per_step"], {k: (round(v["ms_per_step"], 4), round(v["value"]/1e6, 1)) for k, v in dp.items() if isinstance(v, dict)})
d = last("r2_12_weak_n8")
if d: print("e2e N=8:", d["e2e"])
for n in (1, 2, 4, 8):
    d = last(f"r2_12_vlm_n{n}")
    if d: print(f"vlm N={n}: {d['value']/1e6:.2f} M samples/s {d['ms_per_step']:.3f} ms")
PY
