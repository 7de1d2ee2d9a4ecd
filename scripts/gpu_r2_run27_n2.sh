#!/bin/bash
# Round 2, GPU call 27 (TWO GPUs): the data-parallel paths after the pool-backward change (bias sums formed by the tail from
# the cross-rank-summed column sums): dp_check for fused / nccl / peer, the 2-GPU bench, the reference arm under torchrun.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/dp_check.py > gpurun_out/r2_27_dp_check.log 2>&1
echo "dp_check exit $?"; grep -E "^\{|Error|error|differs" gpurun_out/r2_27_dp_check.log | cut -c1-600 | tail -12
grep -E "^\{" gpurun_out/r2_27_dp_check.log > gpurun_out/r2_27_dp_check_n2.jsonl
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_27_bench_n2.json 2> gpurun_out/r2_27_bench_n2.err
echo "n2 exit $?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_27_bench_reference_n2.json 2> gpurun_out/r2_27_bench_reference_n2.err
echo "reference n2 exit $?"; cut -c1-200 gpurun_out/r2_27_bench_reference_n2.json
python - <<'PY'
import json
for f in ("r2_27_bench_n2",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        dp = d["data_parallel"]
        print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 4), d.get("e2e"), {k: (v if not isinstance(v, dict) else str(v)[:80]) for k, v in dp.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 3 gpurun_out/r2_27_bench_n2.err
