#!/bin/bash
# Round-2 pass f (2 GPUs): decode-attention split sweep (GPU 0), the 2-GPU parity tests (exchange, sharded optimizer),
# and the bench under torchrun at N = 2 (carries the ZeRO-2 fine-tune sub-record).
mkdir -p gpurun_out
timeout 120 python tools/decode_attn_bench.py > gpurun_out/r2f_decode_attn_sweep.json 2> gpurun_out/r2f_decode_attn_sweep.err
cat gpurun_out/r2f_decode_attn_sweep.json; tail -3 gpurun_out/r2f_decode_attn_sweep.err
TRAIN=0 BATCH=128 STEPS=2 BENCH_TIMEOUT=500 bash tools/gpu_dist.sh 2 r2f_dist tests
