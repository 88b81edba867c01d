#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_preprocess.py -m gpu -q --timeout 200 -rfE -s 2>&1 | tail -30 > gpurun_out/r2p_pytest_preprocess.log
echo "rc=${PIPESTATUS[0]}"; grep -E "nvjpeg|preprocess\(|passed|failed|Error|assert" gpurun_out/r2p_pytest_preprocess.log | tail -20
