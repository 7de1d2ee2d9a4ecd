#!/bin/bash
# Round 2, GPU call 28 (EIGHT GPUs, short): the final build at N = 8 -- dp_check, weak scaling, strong scaling.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/dp_check.py > gpurun_out/r2_28_dp_check.log 2>&1
echo "dp_check exit $?"; grep -E "^\{" gpurun_out/r2_28_dp_check.log > gpurun_out/r2_28_dp_check_n8.jsonl; cut -c1-330 gpurun_out/r2_28_dp_check_n8.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_28_weak_n8.json 2> gpurun_out/r2_28_weak_n8.err
echo "weak n8 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --global-batch 65536 > gpurun_out/r2_28_strong_n8.json 2> gpurun_out/r2_28_strong_n8.err
echo "strong n8 exit $?"
python - <<'PY'
import json
for f in ("r2_28_weak_n8", "r2_28_strong_n8"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        dp = d["data_parallel"]
        print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 4), dp.get("collective"), dp.get("rank_spread"))
    except Exception as e:
        print(f, "unreadable", e)
PY
