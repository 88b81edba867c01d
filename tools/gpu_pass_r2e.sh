#!/bin/bash
TAG=r2e2
mkdir -p gpurun_out
B200_GEMM_CARVEOUT=1 timeout 200 python tools/corun_bench.py > gpurun_out/${TAG}_corun.json 2> gpurun_out/${TAG}_corun.err
B200_GEMM_CARVEOUT=1 B200_ATTN_CARVEOUT=1 timeout 200 python tools/corun_bench.py >> gpurun_out/${TAG}_corun.json 2>> gpurun_out/${TAG}_corun.err
cat gpurun_out/${TAG}_corun.json; tail -5 gpurun_out/${TAG}_corun.err
