#!/bin/bash
# First GPU pass of a round (1 GPU): everything that was written without hardware access gets run first -- each step
# under its own timeout, reported separately --, then the numbers that DESIGN.md / profiles/README.md quote are
# re-measured. About 30-40 minutes of box time in all (worst case, every timeout hit: 2 h); steps are independent, so
# the script can be cut after any of them. Output in gpurun_out/<tag>_*.
#   gpurun --timeout 3000 -- bash tools/gpu_first_pass.sh r2
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt
# 1. tests that have never run on hardware: short leash, reported separately
timeout 240 python -m pytest tests/test_gpu_zzz_baseline_configs.py tests/test_gpu_zz_pointcloud.py \
  tests/test_gpu_zz_train_extras.py -m gpu -q --timeout 200 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_new.log
echo "new tests rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/${TAG}_pytest_new.log
# 1b. the decode step's opt-in switches (programmatic dependent launch, tile widths 96/160/224; never run on hardware)
B200_TEST_SWITCHES=1 timeout 300 python -m pytest tests/test_gpu_zzzz_switches.py -m gpu -q --timeout 250 2>&1 | tail -8 \
  > gpurun_out/${TAG}_pytest_pdl.log
echo "pdl tests rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/${TAG}_pytest_pdl.log
# 1c. decode-step A/B of the switches at the real 7B size (only meaningful if 1b passed)
timeout 400 python tools/decode_bench.py --batch 128 > gpurun_out/${TAG}_decode_bench.json 2> gpurun_out/${TAG}_decode_bench.err
echo "decode_bench rc=$?"; cat gpurun_out/${TAG}_decode_bench.json; tail -2 gpurun_out/${TAG}_decode_bench.err
# 2. the whole suite + smoke
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
# 3. point-cloud branch: timing (tile / tap-split selection of the gather-GEMM has not been timed yet) + launch list
for N in 20000 60000 150000; do
  timeout 90 python tools/pc_bench.py --points $N --clouds 2 >> gpurun_out/${TAG}_pc_bench.json 2>> gpurun_out/${TAG}_pc_bench.err
done
cat gpurun_out/${TAG}_pc_bench.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_pc_launches.csv \
  python tools/pc_bench.py --points 60000 --clouds 2 --iters 1 > /dev/null 2>&1
echo "ncu pc list rc=$?"
# 4. headline bench (config 2, default batch)
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
# 5. the same bench with programmatic dependent launch (only meaningful if 1b passed) and at a larger decode batch
#    (weights amortised over more samples: KV cache 0.56 GB per sample, 192 samples = 108 GB + 14 GB of weights)
timeout 900 python bench.py --steps 3 --warmup 3 --pdl --no-cpu-baseline > gpurun_out/${TAG}_bench_pdl.json 2> gpurun_out/${TAG}_bench_pdl.err
echo "bench --pdl rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_pdl.json
timeout 900 python bench.py --steps 3 --warmup 3 --decode-tiles 2 --pdl --no-cpu-baseline > gpurun_out/${TAG}_bench_tiles.json 2> gpurun_out/${TAG}_bench_tiles.err
echo "bench --decode-tiles rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_tiles.json
timeout 900 python bench.py --steps 2 --warmup 3 --batch 192 --no-cpu-baseline > gpurun_out/${TAG}_bench_b192.json 2> gpurun_out/${TAG}_bench_b192.err
echo "bench b192 rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_b192.json
# 256 samples per GPU: KV cache 146 GB + 14 GB of weights + ~10 GB of workspaces -- fits only if the allocator has
# little slack; a CUDA out-of-memory error here is a Python exception, not a dead box
timeout 900 python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_bench_b256.json 2> gpurun_out/${TAG}_bench_b256.err
echo "bench b256 rc=$?"; tail -c 1200 gpurun_out/${TAG}_bench_b256.json; tail -2 gpurun_out/${TAG}_bench_b256.err
# 6. fine-tune step at 7B, LoRA vs QLoRA (NF4 round trip of the frozen base, LoRA dropout 0 in both: timing only)
for EXTRA in "" "--nf4"; do
  timeout 600 python tools/train_bench.py --lora-r 128 $EXTRA >> gpurun_out/${TAG}_train_lora.json 2>> gpurun_out/${TAG}_train_lora.err
  echo "train_bench --lora-r 128 $EXTRA rc=$?"
done
cat gpurun_out/${TAG}_train_lora.json; tail -3 gpurun_out/${TAG}_train_lora.err
