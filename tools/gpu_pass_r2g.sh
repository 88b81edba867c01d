#!/bin/bash
# Round-2 pass g (2 GPUs): 2-GPU parity tests on the flat ZeRO-2 step, fine-tune step at N = 2 (ZeRO-2), serving tests.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_zzz_serving.py -m gpu -q --timeout 280 -rfE 2>&1 | tail -40 > gpurun_out/r2g_pytest.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -25 gpurun_out/r2g_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  tools/train_bench.py --zero 2 --steps 3 > gpurun_out/r2g_train_z2_n2.json 2> gpurun_out/r2g_train_z2_n2.err
echo "train_bench zero=2 N=2 rc=$?"; tail -c 1500 gpurun_out/r2g_train_z2_n2.json; tail -3 gpurun_out/r2g_train_z2_n2.err
