"""Multi-GPU sharding of the dmrecon path (DESIGN.md "Multi-GPU").

Reference views are the independent units of the path: every mvs::DMRecon reads only its own image, the pyramids of
its <= globalVSMax neighbours and the bundle, and writes only its own maps (apps/dmrecon/dmrecon.cc:285-318 runs them in
any order on OpenMP threads).  So reference views are block-sharded over ranks with NO collective on the data path.
The single exchange is the INPUT of the neighbours: each rank owns (decodes, uploads) the images of its shard; the images a
rank needs from other shards - the selected neighbours of its reference views, known from the host-side global view
selection - arrive either through one all-gather of all level-0 images (all_gather_images) or, cheaper when a rank needs
only part of the scene, through point-to-point sends of exactly the needed views (exchange_needed_images); each GPU then
builds the pyramids it needs locally.
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


def exchange_needed_images(local: torch.Tensor, owned: List[int], needed: List[int], n_views: int, rank: int, world: int):
    """local: [len(owned), H, W, 3] uint8 - the images of this rank's views; needed: the view ids this rank needs (its reference
    views and their selected neighbours).  Every rank publishes its needed set (one small all-gather of a bit mask), then each
    owner sends exactly the requested images (batched isend / irecv: NCCL over NVLink on GPUs, gloo in the CPU tests).
    Returns {view id: [H, W, 3] tensor} for every needed view and the number of image bytes received."""
    own_pos = {v: k for k, v in enumerate(owned)}
    if world == 1:
        return {v: local[own_pos[v]] for v in needed}, 0
    mask = torch.zeros(n_views, dtype=torch.uint8, device=local.device)
    if needed:
        mask[torch.as_tensor(sorted(needed), dtype=torch.long, device=local.device)] = 1
    masks = torch.empty(world * n_views, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(masks, mask)
    masks = masks.view(world, n_views).cpu().numpy()
    recv_ids = [v for v in sorted(needed) if v not in own_pos]
    recv_buf = torch.empty((len(recv_ids),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    ops = []
    for q in range(world):                      # sends: peers ascending, views ascending (the receivers post the same order)
        if q == rank:
            continue
        for v in owned:
            if masks[q][v]:
                ops.append(dist.P2POp(dist.isend, local[own_pos[v]], q))
    for k, v in enumerate(recv_ids):            # recv_ids ascending = owners ascending, views ascending per owner
        ops.append(dist.P2POp(dist.irecv, recv_buf[k], owner_of(v, n_views, world)))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    out = {v: local[own_pos[v]] for v in needed if v in own_pos}
    out.update({v: recv_buf[k] for k, v in enumerate(recv_ids)})
    return out, int(recv_buf.numel())
