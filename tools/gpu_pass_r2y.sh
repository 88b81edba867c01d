#!/bin/bash
# Last check of the final tree (1 GPU): the GPU suite as the driver runs it + smoke.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -12 > gpurun_out/r2y_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/r2y_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
