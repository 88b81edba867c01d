#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 280 -rfE 2>&1 | tail -10
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  tools/train_bench.py --zero 2 --steps 4 --trace > gpurun_out/r2h2_train_z2_n2_trace.json 2> gpurun_out/r2h2_train_z2_n2_trace.err
echo "n2 rc=$?"; tail -c 2500 gpurun_out/r2h2_train_z2_n2_trace.json; tail -3 gpurun_out/r2h2_train_z2_n2_trace.err
