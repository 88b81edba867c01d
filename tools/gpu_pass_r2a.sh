#!/bin/bash
# Round-2 first GPU pass (1 GPU): the whole suite WITHOUT -x (every failure listed), the opt-in switch tests, the
# decode-step A/B of the switches, smoke, the headline bench (default batch, and 192 samples). Output: gpurun_out/r2a_*.
#   gpurun --timeout 1500 -- bash tools/gpu_pass_r2a.sh
TAG=r2a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -rfE 2>&1 | tail -120 > gpurun_out/${TAG}_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -40 gpurun_out/${TAG}_pytest_gpu.log
B200_TEST_SWITCHES=1 timeout 300 python -m pytest tests/test_gpu_zzzz_switches.py -m gpu -q --timeout 250 -rfE 2>&1 | tail -40 \
  > gpurun_out/${TAG}_pytest_switches.log
echo "switch tests rc=${PIPESTATUS[0]}"; tail -12 gpurun_out/${TAG}_pytest_switches.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python tools/decode_bench.py --batch 128 > gpurun_out/${TAG}_decode_bench.json 2> gpurun_out/${TAG}_decode_bench.err
echo "decode_bench rc=$?"; cat gpurun_out/${TAG}_decode_bench.json; tail -2 gpurun_out/${TAG}_decode_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 2 --warmup 3 --batch 192 --no-cpu-baseline > gpurun_out/${TAG}_bench_b192.json 2> gpurun_out/${TAG}_bench_b192.err
echo "bench b192 rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_b192.json; tail -3 gpurun_out/${TAG}_bench_b192.err
for N in 20000 60000 150000; do
  timeout 90 python tools/pc_bench.py --points $N --clouds 2 >> gpurun_out/${TAG}_pc_bench.json 2>> gpurun_out/${TAG}_pc_bench.err
done
cat gpurun_out/${TAG}_pc_bench.json
