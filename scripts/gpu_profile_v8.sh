#!/bin/bash
# GPU box: v8 evidence -- ncu launch list of the bench command, `--set full` of the cta_group::2 GEMMs and the fused tails,
# bench lines of BASELINE configs 2-4 and the reference arm.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v8.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_v8.log 2>&1
echo "ncu list exit $?" >> gpurun_out/ncu_bench_v8.log
timeout 600 ncu --set full --clock-control none -k regex:"gemm_tc2_kernel|ola_loss_kernel|finalize_norm_kernel|adam_kernel" \
    --launch-skip 16 --launch-count 8 -f -o gpurun_out/v8_gemm_tails python scripts/prof_step.py 3 > gpurun_out/ncu_full_v8.log 2>&1
echo "ncu full exit $?" >> gpurun_out/ncu_full_v8.log
for w in 2 3 4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ls -la gpurun_out | tail -n 15
tail -n 3 gpurun_out/ncu_full_v8.log gpurun_out/ncu_bench_v8.log
python -c "
import json
for f in ['bench','bench_w2','bench_w3','bench_w4','bench_reference']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['ms_per_step'],4), d['value'], d.get('e2e',{}).get('value'), d.get('roofline',{}).get('stages_ms'))
    except Exception as e: print(f, 'ERR', e)
"
