#!/bin/bash
# multi-GPU: weak-scaling bench over NCCL at N = 2 (or $1), plus the reference arm launched the same way
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N exit $?" >> gpurun_out/bench_n$N.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1_same_box.err
tail -c 400 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
