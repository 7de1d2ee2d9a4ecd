#!/bin/bash
# Round 2, GPU call 11 (ONE box): dry run of what the 8-GPU call will run -- the sweep with parity, the VLM step -- plus tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2_11_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_11_tests.log; tail -3 gpurun_out/r2_11_tests.log
timeout 600 python tests/sweep_parity.py > gpurun_out/r2_11_sweep_n1.jsonl 2> gpurun_out/r2_11_sweep_n1.err
echo "sweep exit $?"; grep -v "^{" gpurun_out/r2_11_sweep_n1.jsonl; grep -o '"error": "[^"]*"' gpurun_out/r2_11_sweep_n1.jsonl | head; tail -n 5 gpurun_out/r2_11_sweep_n1.err
timeout 300 python examples/train_vlm.py --steps 10 --launch-list > gpurun_out/r2_11_vlm_n1.json 2> gpurun_out/r2_11_vlm_n1.err
echo "vlm exit $?"; cut -c1-1500 gpurun_out/r2_11_vlm_n1.json; tail -n 5 gpurun_out/r2_11_vlm_n1.err
