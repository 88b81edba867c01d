#!/bin/bash
# Final check of the final tree (1 GPU): the GPU suite exactly as the driver runs it, and smoke.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -15 > gpurun_out/r2r_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -8 gpurun_out/r2r_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
