#!/bin/bash
# GPU box: the reduced-precision mode -- its tests, then bench lines of BASELINE configs 2-4 (extra lines, not the headline).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_precision.py -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_precision.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_precision.log
for w in 2 4 3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
  echo "bench w$w exit $?" >> gpurun_out/bench_w$w.err
done
ST_TF32_AE_FFMA2=1 timeout 600 python bench.py --workload 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_w2_ffma2ae.json 2> gpurun_out/bench_w2_ffma2ae.err
tail -15 gpurun_out/pytest_precision.log
for f in gpurun_out/bench_w*.json; do echo "== $f"; head -c 1800 $f; echo; done
tail -3 gpurun_out/bench_w*.err
