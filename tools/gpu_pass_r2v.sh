#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/train_bench.py --steps 5 --warmup 2 > gpurun_out/r2v_train_n1.json 2> gpurun_out/r2v_train_n1.err
echo "rc=$?"; cat gpurun_out/r2v_train_n1.json; tail -2 gpurun_out/r2v_train_n1.err
