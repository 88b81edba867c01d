#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_zzz_qlora.py -m gpu -q --timeout 60 -rfE -s -k "fused_nf4" 2>&1 | tail -30 > gpurun_out/r2x_nf4.log
echo "rc=${PIPESTATUS[0]}"; grep -E "fused nf4|passed|failed|Error|assert|Timeout" gpurun_out/r2x_nf4.log | tail -20
