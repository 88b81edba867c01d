#!/bin/bash
# fused RoPE + KV append: correctness (switch tests incl. the oracle comparison at 384 CTAs) and the per-step A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zzzz_switches.py tests/test_gpu_e2e.py -m gpu -q --timeout 250 -rfE 2>&1 | tail -12
timeout 300 python tools/decode_bench.py --batch 128 --arms tiles1+nofuse,tiles1,tiles1+nofuse,tiles1 > gpurun_out/r2t_decode_bench_fused_rope.json 2> gpurun_out/r2t.err
echo "rc=$?"; cat gpurun_out/r2t_decode_bench_fused_rope.json; tail -2 gpurun_out/r2t.err
