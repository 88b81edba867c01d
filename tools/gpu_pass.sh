#!/bin/bash
# One GPU pass: parity tests, smoke, the full-config bench, an ncu launch list and ncu --set full captures of the
# dominant kernels. Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
#   usage: tools/gpu_pass.sh [tag] [bench batches, comma separated] [stages: newtests,tests,skinny,bench,list,full]
TAG=${1:-r1}
BATCHES=${2:-64}
STAGES=${3:-tests,bench,list,full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt
(lscpu | head -20; nproc; free -g) > gpurun_out/${TAG}_host.txt 2>&1

if [[ $STAGES == *newtests* ]]; then
  # kernels that have never run on hardware: short leash
  timeout 150 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py -m gpu -x -q -k "${NEWTESTS:-skinny}" --timeout 120 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_new.log
  echo "new tests rc=${PIPESTATUS[0]}"; tail -15 gpurun_out/${TAG}_pytest_new.log
fi

if [[ $STAGES == *,tests* || $STAGES == tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_gpu.log
  echo "pytest rc=${PIPESTATUS[0]}"; tail -8 gpurun_out/${TAG}_pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
fi

if [[ $STAGES == *skinny* ]]; then
  timeout 300 python tools/gpu_gemm_check.py skinny > gpurun_out/${TAG}_skinny_perf.log 2>&1
  echo "skinny perf rc=$?"; grep -E "best-variant|splits=0|tiled bn=|cuBLAS" gpurun_out/${TAG}_skinny_perf.log | tail -40
fi

if [[ $STAGES == *attn* ]]; then
  timeout 200 python tools/gpu_attn_check.py > gpurun_out/${TAG}_attn_perf.log 2>&1
  echo "attn perf rc=$?"; cat gpurun_out/${TAG}_attn_perf.log | tail -20
fi

if [[ $STAGES == *bench* ]]; then
  for B in ${BATCHES//,/ }; do
    extra=""; [[ $B != ${BATCHES%%,*} ]] && extra="--no-cpu-baseline"
    timeout 1500 python bench.py --batch $B --steps 3 --warmup 3 $extra > gpurun_out/${TAG}_bench_b$B.json 2> gpurun_out/${TAG}_bench_b$B.err
    echo "bench B=$B rc=$?"; tail -c 5000 gpurun_out/${TAG}_bench_b$B.json; tail -5 gpurun_out/${TAG}_bench_b$B.err
  done
fi

NCU_CMD="python bench.py --batch 16 --new-tokens 4 --steps 1 --warmup 0 --no-cpu-baseline --no-roofline --no-e2e --no-train --no-configs"
if [[ $STAGES == *list* ]]; then
  # launch list of one whole step (all 32 layers, eager decode loop so that every launch is visible); only this
  # library's kernels (namespace b200) so that torch's weight-initialisation kernels do not eat the launch budget
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b200:: \
    -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv $NCU_CMD --no-graph > gpurun_out/${TAG}_ncu_list.log 2>&1
  echo "ncu list rc=$?"; wc -l gpurun_out/${TAG}_launches.csv
fi

if [[ $STAGES == *full* ]]; then
  # --set full on the kernels that carry the step; 2 decoder layers are enough to reach every kernel shape
  if [[ -n "$NCU_ONLY_FLASH" ]]; then
    SPECS=("flash128 flash_attn_tc_kernel.*128 2 2" "flash64 flash_attn_tc_kernel.*64 2 1")
  else
    SPECS=("gemm256vit gemm_bf16_tn_kernel.*256 40 4" "gemm256llm gemm_bf16_tn_kernel.*256 103 5" \
           "flash128 flash_attn_tc_kernel.*128 2 2" "flash64 flash_attn_tc_kernel.*64 2 1" \
           "decattn decode_attn_kernel 2 2" "gemmskinny gemm_skinny_kernel 2 5")
  fi
  for spec in "${SPECS[@]}"; do
    read name pat skip cnt <<< "$spec"
    timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$pat --launch-skip $skip -c $cnt -f -o gpurun_out/${TAG}_${name} $NCU_CMD --layers 2 \
      > gpurun_out/${TAG}_ncu_${name}.log 2>&1
    echo "ncu $name rc=$? $(grep -c Profiling gpurun_out/${TAG}_ncu_${name}.log) kernels"
  done
fi
ls -la gpurun_out | tail -40
