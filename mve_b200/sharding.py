"""Multi-GPU sharding of the dmrecon path (DESIGN.md "Multi-GPU").

Reference views are the independent units of the path: every mvs::DMRecon reads only its own image, the pyramids of
its <= globalVSMax neighbours and the bundle, and writes only its own maps (apps/dmrecon/dmrecon.cc:285-318 runs them in
any order on OpenMP threads).  So reference views are block-sharded over ranks with NO collective on the data path.
The single exchange is the INPUT of the neighbours: each rank owns (decodes, uploads) the images of its shard and one
all-gather of the uint8 level-0 images gives every GPU every view; each GPU then builds the pyramids it needs locally.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def owned_views(n_views: int, rank: int, world: int) -> List[int]:
    """Contiguous block of view ids owned by `rank` (block sizes differ by at most one)."""
    base, rem = divmod(n_views, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return list(range(lo, hi))


def owner_of(view: int, n_views: int, world: int) -> int:
    base, rem = divmod(n_views, world)
    cut = rem * (base + 1)
    return view // (base + 1) if view < cut else rem + (view - cut) // max(base, 1)


def all_gather_images(local: torch.Tensor, world: int) -> torch.Tensor:
    """local: [n_owned, H, W, 3] uint8 on this rank's device -> [n_views, H, W, 3] in view-id order on every rank.
    Shards may differ in size by one view (padded for the collective, trimmed afterwards)."""
    if world == 1:
        return local
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)
