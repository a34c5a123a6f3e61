"""Data-parallel plumbing.  The reference has no working multi-GPU path (a disabled
nn.DataParallel stub, train.py:260-263); windows are independent, so the batch dimension shards across ranks with
ONE exchange per step: a sum-allreduce of the gradients, scaled by 1/world before the L1 clip (the clip is a function
of the averaged front-end gradients) and Adam (replicated and deterministic, so replicas stay identical).

Everything here is host-side logic on torch tensors and works with any backend: NCCL over NVLink on the GPU box
(one process per GPU), gloo on CPU in the tests."""
import torch
import torch.distributed as dist


class FlatBuffer:
    """One contiguous fp32 buffer with a 16-byte-aligned slot per tensor, so a single collective covers all of them."""

    def __init__(self, shapes, device):
        self.offsets, total = [], 0
        for s in shapes:
            n = 1
            for d in s:
                n *= int(d)
            self.offsets.append(total)
            total += (n + 3) // 4 * 4
        self.flat = torch.zeros(total, device=device, dtype=torch.float32)
        self.views = []
        for o, s in zip(self.offsets, shapes):
            n = 1
            for d in s:
                n *= int(d)
            self.views.append(self.flat[o:o + n].view(tuple(s)))


def nccl_env_defaults(world):
    """NCCL algorithm / protocol for the step's one collective (8.5 MB allreduce), to be called BEFORE the process group is
    created (NCCL reads its environment when the communicator is built); anything the user has set wins.  Measured on B200 /
    NVSwitch boxes with bench.py (ms/step, default selection vs ring + LL128): 2 GPUs 0.782 vs 0.773, 4 GPUs 0.785 vs 0.792,
    8 GPUs 0.815-0.826 vs 0.804-0.808 (tree + LL128 0.830, ring + LL 0.836, ring + Simple 0.920; NVLS-only is refused for the
    float64 control reductions).  Per-collective syntax (NCCL >= 2.24), so broadcast and barrier keep their defaults."""
    import os
    if "NCCL_ALGO" in os.environ or "NCCL_PROTO" in os.environ:
        return
    if world >= 8 or world == 2:
        os.environ["NCCL_ALGO"] = "allreduce:ring"
        os.environ["NCCL_PROTO"] = "allreduce:LL128"


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_sum_(flat, group=None):
    """In-place sum over ranks (no-op for a single process).  Returns the factor the consumer must scale by (1/world)."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w


class GradReducer:
    """Sum-allreduce of the flat gradient buffer in the pieces that actually carry information, overlapped with the backward.

    Layout of the flat buffer (state_dict order): analysis real | analysis imag | synthesis real | synthesis imag | the two
    autoencoders.  Rows >= F of the analysis tensors never receive gradient (cls_fe_dft.py:55-56 slices those bins off), so
    only their first F rows travel: 12.6 MB instead of 16.8 MB at N = 1024.  The synthesis pair is final after the first
    half of the backward (`Engine.backward(part="begin")`), so `start_synthesis()` launches its allreduce (8.4 MB) while the
    autoencoder and analysis gradients are still being computed; `finish()` reduces the rest (4.3 MB on the critical path)
    and waits for everything.  Works with any backend (NCCL on the GPU box, gloo on CPU in the tests)."""

    def __init__(self, fb: FlatBuffer, shapes, live_rows, group=None):
        self.flat, self.group = fb.flat, group
        n = [1] * len(shapes)
        for i, sh in enumerate(shapes):
            for d in sh:
                n[i] *= int(d)
        row = n[0] // int(shapes[0][0])
        o = fb.offsets
        self.analysis = [self.flat[o[0]:o[0] + live_rows * row], self.flat[o[1]:o[1] + live_rows * row]]
        self.synthesis = self.flat[o[2]:o[3] + n[3]]
        self.rest = self.flat[o[4]:]
        self.pending = []

    def start_synthesis(self):
        if world_size(self.group) > 1:
            self.pending.append(dist.all_reduce(self.synthesis, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Returns the factor the consumer must scale by (1/world)."""
        w = world_size(self.group)
        if w > 1:
            for t in self.analysis + [self.rest]:
                self.pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for work in self.pending:
                work.wait()
            self.pending = []
        return 1.0 / w


def pack_payload(views, F):
    """Host-side statement of the packed exchange payload (the CUDA path builds it with st_pack_grads): the first F rows of the
    four DFT tensors -- analysis rows >= F never receive gradient, the synthesis pair is Hermitian -- and the autoencoder
    tensors, concatenated.  Works on any device; used by the CPU tests of the exchange logic."""
    parts = [v.reshape(v.shape[0], -1)[:F].reshape(-1) for v in views[:4]] + [v.reshape(-1) for v in views[4:]]
    return torch.cat(parts)


def unpack_payload(packed, views, F):
    """Inverse of pack_payload: scatters the (reduced) payload back and restores the mirrored synthesis rows
    (g_r[N-k] = g_r[k], g_i[N-k] = -g_i[k], cls_fe_dft.py:109-110).  Analysis rows >= F are left untouched (they are zero)."""
    o = 0
    for t, v in enumerate(views[:4]):
        m = v.reshape(v.shape[0], -1)
        N = m.shape[0]
        m[:F].copy_(packed[o:o + F * m.shape[1]].view(F, -1))
        o += F * m.shape[1]
        if t >= 2:
            mirror = m[1:N // 2].flip(0)
            m[N // 2 + 1:].copy_(mirror if t == 2 else -mirror)
    for v in views[4:]:
        v.reshape(-1).copy_(packed[o:o + v.numel()])
        o += v.numel()


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items windows for `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(tensors, src=0, group=None):
    """Make every replica start from rank `src`'s parameters."""
    if world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)
