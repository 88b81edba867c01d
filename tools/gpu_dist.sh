#!/bin/bash
# N-GPU pass (run with `gpurun --gpus N -- bash tools/gpu_dist.sh N [tag] [tests|notests]`): 2-rank parity tests of the
# sharded path, then the bench under torchrun exactly as the driver launches it. Every step has a hard timeout: a
# multi-GPU hang is charged N x.
N=${1:-2}
TAG=${2:-r1_dist}
MODE=${3:-tests}
mkdir -p gpurun_out
if [[ $MODE == tests ]]; then
  timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q --timeout 280 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
  echo "dist pytest rc=${PIPESTATUS[0]}"; tail -12 gpurun_out/${TAG}_pytest.log
fi
timeout ${BENCH_TIMEOUT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 --batch ${BATCH:-64} --new-tokens ${NEWTOK:-256} \
  > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 3500 gpurun_out/${TAG}_bench_n$N.json; tail -8 gpurun_out/${TAG}_bench_n$N.err
# fine-tune step, data parallel (BASELINE configs[4]): replicated state vs ZeRO-2 (sharded state + bf16 reduce-scatter);
# TRAIN=0 skips it. Full fine-tuning holds 185 GB on one GPU unsharded, so the replicated arm runs the LoRA recipe.
if [[ ${TRAIN:-1} == 1 ]]; then
  for Z in 0 2; do
    EXTRA=""; [[ $Z == 0 ]] && EXTRA="--lora-r 128"
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
      tools/train_bench.py --zero $Z $EXTRA > gpurun_out/${TAG}_train_z${Z}_n$N.json 2> gpurun_out/${TAG}_train_z${Z}_n$N.err
    echo "train_bench zero=$Z N=$N rc=$?"; tail -c 1200 gpurun_out/${TAG}_train_z${Z}_n$N.json; tail -3 gpurun_out/${TAG}_train_z${Z}_n$N.err
  done
fi
