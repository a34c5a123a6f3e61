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


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_sum_(flat, group=None):
    """In-place sum over ranks (no-op for a single process).  Returns the factor the consumer must scale by (1/world)."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w


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
