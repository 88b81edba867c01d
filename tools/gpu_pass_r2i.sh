#!/bin/bash
# Round-2 pass i (1 GPU): ncu launch list of one whole step + ncu --set full of the kernels that carry it (same recipe
# as round 1, tools/gpu_pass.sh), then the bench as the driver runs it but shorter.
bash tools/gpu_pass.sh r2 128 list,full 2>&1 | tail -25
timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/r2i_bench.json; tail -3 gpurun_out/r2i_bench.err
