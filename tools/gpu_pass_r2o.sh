#!/bin/bash
# ncu --set full of the final decode-attention kernel (register-bounded for 4 CTAs per SM) + its standalone rate.
mkdir -p gpurun_out
NCU_CMD="python bench.py --batch 16 --new-tokens 4 --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --no-e2e --no-train --no-configs --layers 2"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:decode_attn_kernel \
  --launch-skip 2 -c 2 -f -o gpurun_out/r2o_decattn $NCU_CMD > gpurun_out/r2o_ncu_decattn.log 2>&1
echo "ncu rc=$?"
SPLITS1=1 timeout 100 python tools/decode_attn_bench.py > gpurun_out/r2o_decode_attn.json 2>&1; cat gpurun_out/r2o_decode_attn.json
