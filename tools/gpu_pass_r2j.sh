#!/bin/bash
# Round-2 pass j (1 GPU): decode-attention occupancy A/B, GEMM raster A/B (throughput + DRAM bytes of the prefill GEMMs).
# (The B200_ATTN_OCC knob existed only for this pass: the 6 / 8-CTA builds were slower and were removed afterwards;
#  profiles/r2_decode_attn_occupancy.json keeps the result.)
mkdir -p gpurun_out
for OCC in 4 6 8; do
  SPLITS1=1 B200_ATTN_OCC=$OCC timeout 100 python tools/decode_attn_bench.py >> gpurun_out/r2j_attn_occ.json 2>> gpurun_out/r2j_attn_occ.err
done
cat gpurun_out/r2j_attn_occ.json; tail -3 gpurun_out/r2j_attn_occ.err
for MB in 0 24 40 64; do
  echo "== B200_RASTER_MB=$MB" >> gpurun_out/r2j_gemm_raster.log
  B200_RASTER_MB=$MB timeout 200 python tools/gpu_gemm_check.py perf 2>&1 | grep -E "bn=0|cuBLAS" | grep -E "vit|llama|square" >> gpurun_out/r2j_gemm_raster.log
done
cat gpurun_out/r2j_gemm_raster.log
NCU_CMD="python bench.py --batch 16 --new-tokens 4 --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --no-e2e --no-train --no-configs --layers 2"
for MB in 0 40; do
  B200_RASTER_MB=$MB timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --kernel-name-base demangled -k regex:gemm_bf16_tn_kernel.*256 --launch-skip 103 -c 5 --csv \
    --log-file gpurun_out/r2j_gemm_dram_mb$MB.csv $NCU_CMD > /dev/null 2>&1
  echo "ncu raster $MB rc=$?"
done
