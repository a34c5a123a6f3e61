# 8-GPU box: the data-parallel bench under a few NCCL algorithm / protocol choices (device-timed ms/step, e2e ms/step).
N=${1:-8}
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/nccl_$tag.json 2> gpurun_out/nccl_$tag.err
  python -c "
import json;b=json.loads(open('gpurun_out/nccl_$tag.json').read().strip().splitlines()[-1]);print('$tag','ms',round(b['ms_per_step'],4),'e2e ms',round(b['e2e']['ms_per_step'],4))" 2>&1 | tail -1
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING
grep -i -m 12 "nvls\|algo\|channels" gpurun_out/nccl_default.err | cut -c1-200
run nvls NCCL_ALGO=NVLS
run tree NCCL_ALGO=Tree
run ring_ll128 NCCL_ALGO=Ring NCCL_PROTO=LL128
run ring_simple NCCL_ALGO=Ring NCCL_PROTO=Simple
