for m in whole sliced overlap; do
  ST_DP_EXCHANGE=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2_$m.json 2> gpurun_out/bench_n2_$m.err
  python -c "
import json;b=json.loads(open('gpurun_out/bench_n2_$m.json').read().strip().splitlines()[-1]);print('$m',b['n_gpus'],b['ms_per_step'],b['value'],b['e2e']['ms_per_step'])"
done
