#!/bin/bash
# the driver's multi-rank launch at N=2: default data-parallel mode, graph replay, e2e, exit path
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final_n2.json 2> gpurun_out/final_n2.err
echo "bench n2 exit $?" >> gpurun_out/final_n2.err
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/final_n2_ref.json 2> gpurun_out/final_n2_ref.err
echo "ref n2 exit $?" >> gpurun_out/final_n2_ref.err
python scripts/show_bench.py gpurun_out/final_n2.json 2>/dev/null | head -7; tail -2 gpurun_out/final_n2.err | cut -c1-200; head -c 200 gpurun_out/final_n2_ref.json; tail -1 gpurun_out/final_n2_ref.err
