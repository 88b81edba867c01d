#!/bin/bash
# Round-2 pass n (2 GPUs): 2-GPU parity tests on the final tree + the bench at full size under torchrun (all sub-records).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 280 -rfE 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
echo "bench N=2 rc=$?"; tail -c 1200 gpurun_out/r2n_bench_n2.json; grep "bench" gpurun_out/r2n_bench_n2.err | tail -12
