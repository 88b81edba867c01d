#!/bin/bash
# Round-2 pass k (8 GPUs): the bench exactly as the driver launches it at N = 8 (shorter), incl. the configs[2]/[3] and
# fine-tune sub-records. Hard timeout: an 8-GPU hang is charged 8x.
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err
echo "bench N=8 rc=$?"; tail -c 7000 gpurun_out/r2k_bench_n8.json; tail -5 gpurun_out/r2k_bench_n8.err
