#!/bin/bash
# first GPU pass of a session: parity tests (hang-safe), smoke, a short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 90 --timeout-method=thread 2>&1 | tail -40 > gpurun_out/pytest_ops.log
echo "ops rc=${PIPESTATUS[0]}"; tail -15 gpurun_out/pytest_ops.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 120 --timeout-method=thread 2>&1 | tail -60 > gpurun_out/pytest_e2e.log
echo "e2e rc=${PIPESTATUS[0]}"; tail -25 gpurun_out/pytest_e2e.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 400 python bench.py --batch 8 --new-tokens 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
