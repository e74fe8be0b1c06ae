"""Batch sharding of the hot path over ranks (one process per GPU).

Imagesets are independent (no cross-batch op anywhere in HRNet.py / Evaluator.py), so the path shards by
contiguous batch ranges with NO data-path collective; the only communication is the optional gather of results
(SR images: 589,824 B per 128x128 imageset; scores: 12 B per imageset) to every rank or to rank 0.  Works with any
torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of n_items imagesets: rank r gets [lo, hi); sizes differ by at most one."""
    if n_items < 0 or world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(lrs: torch.Tensor, alphas: torch.Tensor, rank: Optional[int] = None,
                world_size: Optional[int] = None, group=None) -> Tuple[torch.Tensor, torch.Tensor, Tuple[int, int]]:
    """The slice of a global batch this rank owns (views, no copy).  rank / world_size default to those of `group`."""
    rank = dist.get_rank(group) if rank is None else rank
    world_size = dist.get_world_size(group) if world_size is None else world_size
    lo, hi = shard_range(lrs.shape[0], rank, world_size)
    return lrs[lo:hi], alphas[lo:hi], (lo, hi)


def gather_batch(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """all_gather of per-rank result rows (SR images or score rows) back into global batch order.
    Ragged shards (n_items not divisible by the world size) are padded to the largest shard for the collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    width = max(sizes)
    padded = local
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    bufs: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded.contiguous(), group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)


def sharded_forward_and_score(model, lrs, alphas, hrs=None, hr_maps=None, gather_sr: bool = True, group=None):
    """Runs this rank's shard of (lrs, alphas) through `model` (a callable like HRNet) and, when hrs/hr_maps are given,
    through the shifted-cPSNR search; returns (sr, scores, best_xy) in GLOBAL batch order (gathered).  `model` and the
    scorer are injected so that the host logic can be exercised on CPU with stand-ins."""
    n = lrs.shape[0]
    l_lrs, l_alphas, (lo, hi) = shard_batch(lrs, alphas, group=group)
    if hi > lo:
        sr_local = model(l_lrs, l_alphas)
    else:
        # fewer imagesets than ranks: this rank owns nothing, but it still takes part in the collectives below
        sr_local = torch.zeros((0, 1, 3 * lrs.shape[2], 3 * lrs.shape[3]), dtype=torch.float32, device=lrs.device)
    sr = gather_batch(sr_local, n, group) if gather_sr else sr_local
    if hrs is None:
        return sr, None, None
    if hi > lo:
        from .evaluator import shift_cPSNR_argmax
        best, xy, _ = shift_cPSNR_argmax(sr_local[:, 0], hrs[lo:hi], hr_maps[lo:hi], clip_sr=True)
        packed = torch.cat([torch.as_tensor(best).reshape(-1, 1).float(), torch.as_tensor(xy).reshape(-1, 2).float()], 1)
        packed = packed.to(sr_local.device)
    else:
        packed = torch.zeros((0, 3), dtype=torch.float32, device=sr_local.device)
    packed = gather_batch(packed, n, group)
    return sr, packed[:, 0], packed[:, 1:].to(torch.int64)
