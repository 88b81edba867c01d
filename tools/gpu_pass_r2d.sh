#!/bin/bash
# Round-2 pass d (1 GPU): generate_stream correctness, then serial vs pipelined throughput at the real size.
TAG=r2d
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_zzz_grad_accumulation.py -m gpu -q --timeout 300 -rfE 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 900 python tools/overlap_bench.py --batch 128 --batches 4 --chunks "16,96,16;4,24,4;2,12,2;1,6,1" > gpurun_out/${TAG}_overlap.json 2> gpurun_out/${TAG}_overlap.err
echo "overlap rc=$?"; cat gpurun_out/${TAG}_overlap.json; tail -5 gpurun_out/${TAG}_overlap.err
