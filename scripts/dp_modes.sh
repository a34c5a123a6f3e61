# GPU box with N >= 2 GPUs: the data-parallel exchange variants at N ranks (default 2), then the 2-GPU equivalence test.
N=${1:-2}
for m in packed whole; do
  ST_DP_EXCHANGE=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_$m.json 2> gpurun_out/bench_n${N}_$m.err
  python -c "
import json;b=json.loads(open('gpurun_out/bench_n${N}_$m.json').read().strip().splitlines()[-1]);print('$m',b['n_gpus'],b['ms_per_step'],b['value'],b['e2e']['ms_per_step'],b.get('replica_parameters_identical'))"
done
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json;b=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('single',b['n_gpus'],b['ms_per_step'],b['value'],b['e2e']['ms_per_step'])"
python -m pytest tests/test_gpu_dp.py tests/test_gpu_parity.py::test_packed_exchange_payload_roundtrip -x -q 2>&1 | tail -3
