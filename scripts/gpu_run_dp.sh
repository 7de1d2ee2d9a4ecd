#!/bin/bash
# multi-GPU: weak-scaling bench over NCCL at N ranks, overlapped and non-overlapped gradient all-reduce
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for OV in 1 0; do
  AECF_DP_OVERLAP=$OV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$OV \
      bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_n${N}_overlap$OV.json 2> gpurun_out/bench_n${N}_overlap$OV.err
  echo "bench n$N overlap $OV exit $?" >> gpurun_out/bench_n${N}_overlap$OV.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1_same_box.err
for f in gpurun_out/bench_n${N}_overlap1.json gpurun_out/bench_n${N}_overlap0.json gpurun_out/bench_n$N.json gpurun_out/bench_n1_same_box.json; do python scripts/show_bench.py $f 2>/dev/null | head -2; done
tail -3 gpurun_out/bench_n${N}_overlap1.err
