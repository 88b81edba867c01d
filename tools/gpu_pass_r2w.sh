#!/bin/bash
# prefix-KV reuse test first (short leash), then the whole GPU suite as the driver runs it + smoke
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_zzz_serving.py -m gpu -q --timeout 150 -rfE 2>&1 | tail -25
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -12 > gpurun_out/r2w_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/r2w_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
