#!/bin/bash
# Round-2 final pass (1 GPU): the whole GPU suite exactly as the driver runs it (-x), smoke, the bench at full size.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -15 > gpurun_out/r2m_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -8 gpurun_out/r2m_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
echo "bench rc=$?"; tail -c 5000 gpurun_out/r2m_bench.json; grep "bench" gpurun_out/r2m_bench.err | tail
