"""Data-parallel plumbing for the hot path (one process per GPU, torch.distributed).

The path shards by tokens (K1 microbenchmark, mHC layers) or by images (decode / NMS): units are
independent, so the forward needs NO collective.  Training adds exactly one exchange, the DDP-style
all-reduce of the (small) parameter gradients; NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.distributed as dist

PARAM_GRAD_KEYS = ("dphi", "dbias", "dalpha", "dscale")


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `total` units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_param_grads(grads: Dict[str, torch.Tensor], group=None, keys: Sequence[str] = PARAM_GRAD_KEYS,
                          average: bool = False) -> Dict[str, torch.Tensor]:
    """One flat all-reduce of the parameter gradients (what DDP's bucket does for this layer)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return grads
    flat = torch.cat([grads[k].reshape(-1) for k in keys])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    out = dict(grads)
    off = 0
    for k in keys:
        n = grads[k].numel()
        out[k] = flat[off:off + n].view_as(grads[k])
        off += n
    return out


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-side timing is reported as the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_image_shards(local: Sequence, world: int, group=None):
    """Inference shards images across ranks with no collective on the data path; this host-side gather
    of the per-image detection lists is only for a caller that wants them in one place."""
    if not dist.is_initialized() or world == 1:
        return list(local)
    out = [None] * world
    dist.all_gather_object(out, list(local), group=group)
    return [d for part in out for d in part]
