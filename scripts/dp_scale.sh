# GPU box with N GPUs: the bench at 1..N ranks exactly as the driver launches it (packed exchange, the default), one line each.
N=${1:-8}
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python -c "
import json;b=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1]);print('n',b['n_gpus'],'ms',round(b['ms_per_step'],4),'frames/s',round(b['value']/1e9,4),'G  e2e ms',round(b['e2e']['ms_per_step'],4),'replicas identical',b.get('replica_parameters_identical'))"
done
