#!/bin/bash
# GPU box: round-2 evidence -- ncu launch list of the bench command, `--set full` of the TMEM autoencoder kernels and the GEMMs,
# bench lines of BASELINE configs 1-4 and the reference arm.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r02.log 2>&1
echo "ncu list exit $?" >> gpurun_out/ncu_bench_r02.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ae_fwd_tm_kernel|ae_bwd_tm_kernel|ae_track_to_spec_kernel|ae_pack_kernel|ae_grad_reduce" \
    --launch-skip 10 --launch-count 6 -f -o gpurun_out/r02_ae_tm python scripts/prof_step.py 3 > gpurun_out/ncu_full_r02.log 2>&1
echo "ncu full exit $?" >> gpurun_out/ncu_full_r02.log
for w in 2 3 4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -n 3 gpurun_out/ncu_full_r02.log gpurun_out/ncu_bench_r02.log
python -c "
import json
for f in ['bench','bench_w2','bench_w3','bench_w4','bench_reference']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['ms_per_step'],4), d['value'], d.get('e2e',{}).get('value'), d.get('roofline',{}).get('stages_ms'), d.get('oracle_replay',{}).get('max_abs_diff') if d.get('oracle_replay') else None)
    except Exception as e: print(f, 'ERR', e)
"
