"""Multi-GPU plumbing of the hot path: images are independent through all denoising steps
(model/pipeline.py:411-469 has no cross-sample op), so the global batch is sharded over one process
per GPU with NO data-path collective, and the final latents are collected with a single all-gather
(NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, first+count) image indices of `rank`; earlier ranks take the remainder."""
    if global_batch < 0 or not (0 <= rank < world):
        raise ValueError("bad shard request")
    base, rem = divmod(global_batch, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_latents(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """All ranks receive the (global_batch, ...) latents in image order.  Equal shards use one
    all_gather_into_tensor; ragged shards pad to the largest shard first."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_range(global_batch, r, world)[1] for r in range(world)]
    mx = max(counts)
    buf = local
    if local.shape[0] != mx:
        buf = torch.zeros(mx, *local.shape[1:], dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    out = torch.empty(world * mx, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """Data-parallel stage-2 training (train_StorySalon_stage2.py:168-177, 327 under accelerate's DDP): average the
    gradients of the trainable (attn3) parameters over the ranks with ONE all-reduce of a flat fp32 bucket, in place.
    The 80 attn3 tensors are 198 MB of fp32 gradients: ~0.5 ms over NVLink / NVSwitch against a ~145 ms step, so a single
    bucket after backward is enough (nothing to overlap).  Returns the number of bytes reduced (0 on a single rank)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * 4
