#!/bin/bash
# Round 2, GPU call 8 (TWO GPUs): first multi-process run of the in-backward gradient sum (CUDA IPC peer mappings).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_8_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/dp_check.py > gpurun_out/r2_8_dp_check.log 2>&1
echo "dp_check exit $?"; grep -E "^\{|Error|error|differs" gpurun_out/r2_8_dp_check.log | cut -c1-400 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_8_bench_n1.json 2> gpurun_out/r2_8_bench_n1.err
echo "n1 exit $?"; python scripts/show_bench.py gpurun_out/r2_8_bench_n1.json | head -1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_8_bench_n2.json 2> gpurun_out/r2_8_bench_n2.err
echo "n2 exit $?"; python scripts/show_bench.py gpurun_out/r2_8_bench_n2.json | head -1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --dp-study > gpurun_out/r2_8_bench_n2_study.json 2> gpurun_out/r2_8_bench_n2_study.err
echo "n2 study exit $?"
python - <<'PY'
import json
for f in ("r2_8_bench_n2", "r2_8_bench_n2_study"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        dp = d["data_parallel"]
        print(f, d["value"] / 1e6, d["ms_per_step"], {k: (v if not isinstance(v, dict) else (v["ms_per_step"], v["value"] / 1e6)) for k, v in dp.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r2_8_bench_n2.err gpurun_out/r2_8_bench_n2_study.err
