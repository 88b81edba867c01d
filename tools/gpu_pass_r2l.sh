#!/bin/bash
# Round-2 pass l (2 GPUs, reduced sizes): the complete bench path incl. configs[2]/[3] and fine-tune sub-records under
# torchrun, with progress lines on stderr -- validates what the 8-GPU run did not finish.
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 2 --steps 1 --warmup 1 --batch 32 --new-tokens 32 --config-batch 16 --cfg4-batch 16 --config-steps 1 \
  --train-steps 2 --subrecord-timeout 90 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err
echo "bench N=2 rc=$?"; tail -c 4000 gpurun_out/r2l_bench_n2.json; grep "bench" gpurun_out/r2l_bench_n2.err | tail -20
