#!/bin/bash
# bring-up run on the GPU box: all GPU tests without -x, logs into gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"
tail -40 gpurun_out/pytest_gpu.log
