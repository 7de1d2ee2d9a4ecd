#!/bin/bash
# Round 2, GPU call 31 (FOUR GPUs, short): the final build at N = 4 -- weak and strong scaling.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_31_weak_n4.json 2> gpurun_out/r2_31_weak_n4.err
echo "weak n4 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --global-batch 65536 > gpurun_out/r2_31_strong_n4.json 2> gpurun_out/r2_31_strong_n4.err
echo "strong n4 exit $?"
python - <<'PY'
import json
for f in ("r2_31_weak_n4", "r2_31_strong_n4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        dp = d["data_parallel"]
        print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 4), dp.get("collective"), dp.get("rank_spread"))
    except Exception as e:
        print(f, "unreadable", e)
PY
