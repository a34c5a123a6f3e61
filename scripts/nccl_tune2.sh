# 8-GPU box, second sweep: per-collective algorithm selection (NCCL >= 2.24 syntax) so broadcast / barrier keep their defaults.
N=${1:-8}
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/nccl2_$tag.json 2> gpurun_out/nccl2_$tag.err
  python -c "
import json;b=json.loads(open('gpurun_out/nccl2_$tag.json').read().strip().splitlines()[-1]);print('$tag','ms',round(b['ms_per_step'],4),'e2e ms',round(b['e2e']['ms_per_step'],4))" 2>&1 | tail -1
}
run ar_nvls "NCCL_ALGO=allreduce:nvls"
run ar_ring_ll128 "NCCL_ALGO=allreduce:ring" "NCCL_PROTO=allreduce:LL128"
run ar_tree_ll128 "NCCL_ALGO=allreduce:tree" "NCCL_PROTO=allreduce:LL128"
run ar_ring_ll "NCCL_ALGO=allreduce:ring" "NCCL_PROTO=allreduce:LL"
