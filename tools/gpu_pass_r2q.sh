#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/decode_bench.py --batch 128 --arms none,pdl,none,pdl > gpurun_out/r2q_decode_bench_late_pdl.json 2> gpurun_out/r2q_decode_bench.err
echo "rc=$?"; cat gpurun_out/r2q_decode_bench_late_pdl.json; tail -2 gpurun_out/r2q_decode_bench.err
