"""Two-GPU data-parallel equivalence through the CUDA path and NCCL: 2 ranks x B windows == 1 rank x 2B windows
(SURVEY.md section 8e), up to fp32 summation order.  Skipped on boxes with fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, B, steps, out):
    import torch.distributed as dist
    import signaltrain_b200 as st
    from signaltrain_b200 import data, parallel
    from signaltrain_b200.train import FusedTrainer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    torch.manual_seed(218)
    model = st.nn_proc.st_model(1, 4, 4).to(f"cuda:{rank}")
    lr, _ = st.learningrate.get_1cycle_schedule(1e-4, 200000, 1000, 200)
    tr = FusedTrainer(model, lr, distributed=world > 1)
    x, y, k = data.make_pool(B * steps, model.in_chunk_size, model.out_chunk_size, data.Compressor_4c(), seed=7)
    losses = []
    for s in range(steps):
        lo, hi = parallel.shard_range(B, rank, world)
        sl = slice(s * B + lo, s * B + hi)
        dev = lambda a: torch.from_numpy(a[sl]).to(f"cuda:{rank}")
        losses.append(tr.step(dev(x), dev(y), dev(k)).item())
    sd = {n: p.detach().cpu().numpy() for n, p in model.state_dict().items()}
    if rank == 0:
        np.savez(out, losses=np.array(losses), **{n.replace(".", "/"): a for n, a in sd.items()})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_one_rank(tmp_path):
    import torch.multiprocessing as mp
    B, steps = 8, 2
    ctx = mp.get_context("spawn")
    one, two = str(tmp_path / "one.npz"), str(tmp_path / "two.npz")
    p = ctx.Process(target=_run, args=(0, 1, _free_port(), B, steps, one))
    p.start(); p.join(300)
    assert p.exitcode == 0
    port = _free_port()
    ps = [ctx.Process(target=_run, args=(r, 2, port, B, steps, two)) for r in range(2)]
    for q in ps:
        q.start()
    for q in ps:
        q.join(300)
        assert q.exitcode == 0
    a, b = np.load(one), np.load(two)
    # rank 0's loss is the mean over ITS half; the parameters must agree
    for n in a.files:
        if n == "losses":
            continue
        diff = np.abs(a[n] - b[n])
        assert diff.max() <= 2.5e-5, (n, diff.max())
        assert (diff > 3e-6).mean() < 5e-3, (n, (diff > 3e-6).mean())
