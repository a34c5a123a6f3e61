"""World-size-2 gloo test of the data-parallel logic on CPU: two ranks with B windows each, one gradient
allreduce, scaled clip + Adam  ==  one rank with 2B windows (up to fp summation order).  The compute here is the
numpy oracle -- the point is the host-side exchange/scale/update logic of signaltrain_b200.parallel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import st_oracle as O
from signaltrain_b200 import parallel
from tests.helpers import initial_params, load_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, d = load_case("comp4c_c8192_k4_b3")
    P = initial_params(g, d)
    x, y, k = g["step0/x"][:2], g["step0/y"][:2].astype(np.float32), g["step0/knobs"][:2]
    lo, hi = parallel.shard_range(2, rank, world)
    sbf = O.scale_by_freq(d.F)
    _, grads, _ = O.loss_and_grads(d, P, x[lo:hi], y[lo:hi], k[lo:hi], sbf)
    shapes = [s for _, s in O.param_order(d)]
    fb = parallel.FlatBuffer(shapes, "cpu")
    for v, (name, _) in zip(fb.views, O.param_order(d)):
        v.copy_(torch.from_numpy(np.ascontiguousarray(grads[name], dtype=np.float32)))
    # the production exchange: synthesis pair first (overlappable), then live analysis rows + autoencoders; dead rows stay home
    fb2 = parallel.FlatBuffer(shapes, "cpu")
    fb2.flat.copy_(fb.flat)
    red = parallel.GradReducer(fb2, shapes, d.F)
    red.start_synthesis()
    scale2 = red.finish()
    # the packed exchange (default on the GPU path): one collective over 8.5 MB instead of 16.8 MB
    fb3 = parallel.FlatBuffer(shapes, "cpu")
    fb3.flat.copy_(fb.flat)
    packed = parallel.pack_payload(fb3.views, d.F)
    assert packed.numel() == 4 * d.F * d.N + sum(int(np.prod(s)) for s in shapes[4:])
    scale3 = parallel.allreduce_sum_(packed)
    parallel.unpack_payload(packed, fb3.views, d.F)
    scale = parallel.allreduce_sum_(fb.flat)
    assert scale2 == scale and torch.equal(fb.flat, fb2.flat), "sliced allreduce differs from the whole-buffer allreduce"
    assert scale3 == scale
    for a_, b_ in zip(fb.views, fb3.views):   # Hermitian symmetry of the oracle's float64 gradients holds to rounding
        assert (a_ - b_).abs().max().item() <= 1e-6 * max(a_.abs().max().item(), 1e-30), "packed exchange differs from the whole-buffer one"
    avg = {name: v.numpy().astype(np.float64) * scale for v, (name, _) in zip(fb.views, O.param_order(d))}
    total = O.clip_grad_norm_(avg)
    Pn = O.adam_step({n: a.astype(np.float64) for n, a in P.items()}, avg, {}, 1e-4 / 15)
    if rank == 0:
        q.put((total, {n: Pn[n] for n in ("mpaec.aenc.fnn_enc.weight", O.DFT_KEYS[0], O.DFT_KEYS[2])}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_big_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g, d = load_case("comp4c_c8192_k4_b3")
    P = initial_params(g, d)
    sbf = O.scale_by_freq(d.F)
    _, grads, _ = O.loss_and_grads(d, P, g["step0/x"][:2], g["step0/y"][:2].astype(np.float32), g["step0/knobs"][:2], sbf)
    ref_total = O.clip_grad_norm_(grads)
    Pn = O.adam_step({n: a.astype(np.float64) for n, a in P.items()}, grads, {}, 1e-4 / 15)
    assert abs(total - ref_total) / ref_total < 1e-5
    for n, a in got.items():
        np.testing.assert_allclose(a, Pn[n], atol=2e-7)


def test_shard_range_covers_everything():
    for n in (1, 7, 200, 513):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_flat_buffer_alignment():
    fb = parallel.FlatBuffer([(3,), (5, 7), (1024, 1, 1024), (9,)], "cpu")
    assert all(o % 4 == 0 for o in fb.offsets)
    assert [tuple(v.shape) for v in fb.views] == [(3,), (5, 7), (1024, 1, 1024), (9,)]
    fb.views[1].fill_(2.0)
    assert fb.flat[fb.offsets[1]:fb.offsets[1] + 35].eq(2.0).all() and fb.flat[fb.offsets[1] + 35] == 0


def test_nccl_env_defaults_only_fill_what_the_user_left_unset(monkeypatch):
    """parallel.nccl_env_defaults: ring + LL128 for the step's allreduce at the world sizes where it measured faster (2, >= 8),
    NCCL's own choice at 4, and never over something the user exported."""
    from signaltrain_b200 import parallel
    for w, want in ((2, True), (4, False), (8, True), (16, True), (1, False)):
        monkeypatch.delenv("NCCL_ALGO", raising=False)
        monkeypatch.delenv("NCCL_PROTO", raising=False)
        parallel.nccl_env_defaults(w)
        assert (os.environ.get("NCCL_ALGO") == "allreduce:ring") == want, w
        assert (os.environ.get("NCCL_PROTO") == "allreduce:LL128") == want, w
    monkeypatch.setenv("NCCL_ALGO", "Tree")
    monkeypatch.delenv("NCCL_PROTO", raising=False)
    parallel.nccl_env_defaults(8)
    assert os.environ["NCCL_ALGO"] == "Tree" and "NCCL_PROTO" not in os.environ
