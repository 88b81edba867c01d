#!/bin/bash
# Round-2 third GPU pass (1 GPU): whole suite on the current tree, fine-tune step at 7B (stored / recomputed activations,
# per-family profile, LoRA, QLoRA), headline bench with the train sub-record.
TAG=r2c
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rfE -s 2>&1 | tail -80 > gpurun_out/${TAG}_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; grep -E "per-step|full size|passed|failed|FAILED|Error" gpurun_out/${TAG}_pytest_gpu.log | tail -30
timeout 400 python tools/train_bench.py --steps 2 --prof > gpurun_out/${TAG}_train_full.json 2> gpurun_out/${TAG}_train_full.err
echo "train full rc=$?"; cat gpurun_out/${TAG}_train_full.json; tail -3 gpurun_out/${TAG}_train_full.err
timeout 400 python tools/train_bench.py --steps 2 --recompute > gpurun_out/${TAG}_train_recompute.json 2> gpurun_out/${TAG}_train_recompute.err
echo "train recompute rc=$?"; cat gpurun_out/${TAG}_train_recompute.json; tail -3 gpurun_out/${TAG}_train_recompute.err
for EXTRA in "" "--nf4"; do
  timeout 400 python tools/train_bench.py --lora-r 128 --steps 2 $EXTRA >> gpurun_out/${TAG}_train_lora.json 2>> gpurun_out/${TAG}_train_lora.err
  echo "train_bench --lora-r 128 $EXTRA rc=$?"
done
cat gpurun_out/${TAG}_train_lora.json; tail -3 gpurun_out/${TAG}_train_lora.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 4500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
