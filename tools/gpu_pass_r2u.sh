#!/bin/bash
# Final check of the final tree (1 GPU): the GPU suite as the driver runs it, smoke, a short full-size bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -15 > gpurun_out/r2u_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/r2u_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
echo "bench rc=$?"; head -c 600 gpurun_out/r2u_bench.json; echo; grep "bench" gpurun_out/r2u_bench.err | tail -4
