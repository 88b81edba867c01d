#!/bin/bash
# Round-2 second GPU pass (1 GPU): the tests fixed / added after pass r2a (chain-weight exact tokens, NF4, full-size
# properties and the full-size oracle comparisons), and one repetition of the CPU reference arm with all 255 decode steps.
TAG=r2b
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_zzz_qlora.py tests/test_gpu_zzz_baseline_configs.py \
  tests/test_gpu_zzz_full_size_properties.py -m gpu -q --timeout 600 -rfE -s 2>&1 | tail -60 > gpurun_out/${TAG}_pytest.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -30 gpurun_out/${TAG}_pytest.log
free -g | head -2; nproc
timeout 300 python bench.py --impl reference --steps 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
echo "reference arm rc=$?"; cat gpurun_out/${TAG}_reference_arm.json; tail -2 gpurun_out/${TAG}_reference_arm.err
