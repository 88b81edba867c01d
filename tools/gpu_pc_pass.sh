#!/bin/bash
# Short GPU pass for the point-cloud branch: full parity suite + smoke, timing of PointTransformerV3 on realistic
# clouds, ncu launch list and --set full captures of its two heavy kernels. Output in gpurun_out/.
TAG=${1:-r1_pc}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12 > gpurun_out/${TAG}_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
for N in 20000 60000 150000; do
  timeout 90 python tools/pc_bench.py --points $N --clouds 2 >> gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
done
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python tools/pc_bench.py --points 60000 --clouds 2 --iters 1 > /dev/null 2>&1
echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gather_gemm_kernel|patch_attention_kernel' -s 30 -c 10 \
  -f -o gpurun_out/${TAG}_full python tools/pc_bench.py --points 60000 --clouds 2 --iters 1 > /dev/null 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out | tail -8
