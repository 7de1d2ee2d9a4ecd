#!/bin/bash
# N ranks, graph replay: default all-reduce mode, the other mode and no all-reduce, per-rank times (bench.py --dp-study)
N=${1:-8}
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --dp-study > gpurun_out/dpstudy_n$N.json 2> gpurun_out/dpstudy_n$N.err
echo "bench n$N exit $?" >> gpurun_out/dpstudy_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/dpstudy_n$N.json") if l.startswith("{")][-1])
    print("value", d["value"] / 1e6, "ms", d["ms_per_step"]); print(json.dumps(d.get("data_parallel"), indent=1))
except Exception as e:
    print("no result:", e)
PY
tail -4 gpurun_out/dpstudy_n$N.err | cut -c1-300
