#!/bin/bash
# GPU box: the cta_group::2 GEMM variant -- GEMM tests with it forced on, then the bench with and without it.
set -u
mkdir -p gpurun_out
ST_GEMM_PAIR=1 timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_precision.py -k "gemm" -q -s --timeout 200 -p no:cacheprovider > gpurun_out/pytest_pair.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/pytest_pair.log
grep -E "GEMM a_mn|promoted|passed|failed|Error|error|exit" gpurun_out/pytest_pair.log | tail -40
if [[ $rc == 0 ]]; then
  ST_GEMM_PAIR=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -p no:cacheprovider -x -k "forward_vs_oracle or backward_vs" > gpurun_out/pytest_pair_parity.log 2>&1
  echo "parity exit $?" >> gpurun_out/pytest_pair_parity.log
  tail -4 gpurun_out/pytest_pair_parity.log
  for m in 1 0; do
    ST_GEMM_PAIR=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pair$m.json 2> gpurun_out/bench_pair$m.err
    echo "== pair=$m"; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_pair$m.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:v for k,v in d['roofline']['stages_ms'].items() if 'gemm' in k})"
  done
fi
