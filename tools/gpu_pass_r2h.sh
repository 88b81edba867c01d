#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/train_bench.py --steps 2 --trace > gpurun_out/r2h_train_n1_trace.json 2> gpurun_out/r2h_train_n1_trace.err
echo "n1 rc=$?"; cat gpurun_out/r2h_train_n1_trace.json; tail -2 gpurun_out/r2h_train_n1_trace.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  tools/train_bench.py --zero 2 --steps 2 --trace > gpurun_out/r2h_train_z2_n2_trace.json 2> gpurun_out/r2h_train_z2_n2_trace.err
echo "n2 rc=$?"; tail -c 2500 gpurun_out/r2h_train_z2_n2_trace.json; tail -3 gpurun_out/r2h_train_z2_n2_trace.err
