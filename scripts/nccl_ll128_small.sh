for n in 2 4; do
NCCL_ALGO=allreduce:ring NCCL_PROTO=allreduce:LL128 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ll128_n$n.json 2> gpurun_out/ll128_n$n.err
python -c "
import json;b=json.loads(open('gpurun_out/ll128_n$n.json').read().strip().splitlines()[-1]);print('ring_ll128 n',b['n_gpus'],'ms',round(b['ms_per_step'],4),'e2e ms',round(b['e2e']['ms_per_step'],4))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/def_n$n.json 2> gpurun_out/def_n$n.err
python -c "
import json;b=json.loads(open('gpurun_out/def_n$n.json').read().strip().splitlines()[-1]);print('default    n',b['n_gpus'],'ms',round(b['ms_per_step'],4),'e2e ms',round(b['e2e']['ms_per_step'],4))"
done
