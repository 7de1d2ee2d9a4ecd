#!/bin/bash
# weak-scaling bench on one box: N = 8, 4, 1 back to back (run under `gpurun --gpus 8`)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
PORT=29530
for N in 8 4; do
  PORT=$((PORT+1))
  EXTRA="--no-e2e"; [ "$N" = "8" ] && EXTRA=""
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  echo "bench n$N exit $?" >> gpurun_out/scale_n$N.err
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
echo "bench n1 exit $?" >> gpurun_out/scale_n1.err
for N in 8 4 1; do python scripts/show_bench.py gpurun_out/scale_n$N.json 2>/dev/null | head -2; tail -2 gpurun_out/scale_n$N.err; done
