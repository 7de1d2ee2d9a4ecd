#!/bin/bash
# Round 2, GPU call 10 (TWO GPUs): IPC mapped into the local device's context; dp_debug, dp_check, 2-GPU bench.
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    scripts/dp_debug.py > gpurun_out/r2_10_dp_debug.log 2>&1
echo "dp_debug exit $?"; grep -E "^\[rank|Error|error" gpurun_out/r2_10_dp_debug.log | cut -c1-250 | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/dp_check.py > gpurun_out/r2_10_dp_check.log 2>&1
echo "dp_check exit $?"; grep -E "^\{|Error|error|differs" gpurun_out/r2_10_dp_check.log | cut -c1-500 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_10_bench_n1.json 2> gpurun_out/r2_10_bench_n1.err
echo "n1 exit $?"; python scripts/show_bench.py gpurun_out/r2_10_bench_n1.json 2>/dev/null | head -1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_10_bench_n2.json 2> gpurun_out/r2_10_bench_n2.err
echo "n2 exit $?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --dp-study > gpurun_out/r2_10_bench_n2_study.json 2> gpurun_out/r2_10_bench_n2_study.err
echo "n2 study exit $?"
python - <<'PY'
import json
for f in ("r2_10_bench_n2", "r2_10_bench_n2_study"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        dp = d["data_parallel"]
        print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 4), {k: (v if not isinstance(v, dict) else (round(v["ms_per_step"], 4), round(v["value"] / 1e6, 1))) for k, v in dp.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 4 gpurun_out/r2_10_bench_n2.err
