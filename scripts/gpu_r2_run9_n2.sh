#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    scripts/dp_debug.py > gpurun_out/r2_9_dp_debug.log 2>&1
echo "dp_debug exit $?"; grep -E "^\[rank|Error|error" gpurun_out/r2_9_dp_debug.log | cut -c1-300 | head -30
