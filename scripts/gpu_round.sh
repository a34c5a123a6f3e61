#!/bin/bash
# Runs on the GPU box under gpurun: parity tests, smoke, bench, and (optionally) the ncu launch list.
# Usage: scripts/gpu_round.sh [tests|bench|ncu|all]
set -u
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvsmi.csv 2>&1
if [[ $what == tests || $what == all ]]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.log
fi
if [[ $what == bench || $what == all ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err
fi
if [[ $what == ncu || $what == all ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?" >> gpurun_out/ncu_bench.log
fi
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.err 2>/dev/null
cat gpurun_out/bench.json 2>/dev/null | head -c 3000
