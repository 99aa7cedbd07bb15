"""Multi-GPU sharding of the relevancy path: one process per GPU, per-sample units sharded contiguously across
ranks (weights replicated), no data-path collective, ONE all-gather of the final maps.

The reference has no multi-GPU relevancy path (its runs are single-GPU, batch 1); per-sample independence follows
from the objective being the sum of the DIAGONAL logits only (CLIP_explainability.ipynb:156-160)."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n units for `rank`; the first n % world ranks get one extra unit."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_maps(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of per-sample maps [n_local, ...] -> [n_total, ...] in rank order.  Equal shards use a single
    all_gather_into_tensor (NCCL: one ncclAllGather); ragged shards are padded to the largest shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    if world == 1:
        return local
    mx = max(counts)
    tail = local.shape[1:]
    if min(counts) == mx:
        out = torch.empty((world * mx,) + tuple(tail), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = torch.zeros((mx,) + tuple(tail), dtype=local.dtype, device=local.device)
    padded[:counts[rank]] = local
    out = torch.empty((world * mx,) + tuple(tail), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx:r * mx + counts[r]] for r in range(world)], dim=0)


def gather_maps_packed(R_text: torch.Tensor, R_image: torch.Tensor, n_total: int, group=None, async_op: bool = False):
    """THE collective of the path: both maps of every local sample packed into one ``[n_local, ctx*ctx + P]``
    buffer and exchanged with ONE ``all_gather_into_tensor`` (NCCL: one ncclAllGather).  Equal shards only (the bench /
    serving shape; ragged shards go through :func:`all_gather_maps`).  Returns ``(work, full)``: ``full`` is
    ``[n_total, ctx*ctx + P]`` in rank order (split it with :func:`split_packed`); with ``async_op`` the collective runs on
    NCCL's stream and ``work.wait()`` orders the current stream after it - enqueue the next step's forward in between."""
    world = dist.get_world_size(group)
    n_local = R_text.shape[0]
    assert n_local * world == n_total, "gather_maps_packed needs equal shards"
    packed = torch.cat((R_text.reshape(n_local, -1), R_image.reshape(n_local, -1)), dim=1)
    full = torch.empty((n_total, packed.shape[1]), dtype=packed.dtype, device=packed.device)
    work = dist.all_gather_into_tensor(full, packed, group=group, async_op=async_op)
    return work, full


def split_packed(full: torch.Tensor, ctx: int, patches: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """``[n, ctx*ctx + patches]`` -> ``(R_text [n,ctx,ctx], R_image [n,patches])`` (views)."""
    n = full.shape[0]
    return full[:, :ctx * ctx].reshape(n, ctx, ctx), full[:, ctx * ctx:ctx * ctx + patches]


def interpret_sharded(interpret_fn: Callable, images: torch.Tensor, tokens: torch.Tensor, start_layer: int = -1,
                      start_layer_text: int = -1, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every rank holds the same global batch; rank r computes samples shard_range(B, r, world) with
    ``interpret_fn(images, tokens, start_layer, start_layer_text)`` (e.g. ``engine.interpret``) and all ranks
    receive the full ``(R_text [B,ctx,ctx], R_image [B,S-1])``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = tokens.shape[0]
    lo, hi = shard_range(B, rank, world)
    img = images if images.shape[0] == 1 else images[lo:hi]
    if hi > lo:
        rt, ri = interpret_fn(img, tokens[lo:hi], start_layer, start_layer_text)
    else:   # more ranks than samples: contribute an empty shard with the right trailing shape
        rt0, ri0 = interpret_fn(images[:1], tokens[:1], start_layer, start_layer_text)
        rt, ri = rt0[:0], ri0[:0]
    if world == 1:
        return rt, ri
    if B % world == 0:                                   # equal shards: ONE all-gather of the packed maps
        _, full = gather_maps_packed(rt, ri, B, group)
        return split_packed(full, rt.shape[-1], ri.shape[-1])
    return all_gather_maps(rt, B, group), all_gather_maps(ri, B, group)


def map_sharded(unit_fn: Callable[[int, int], torch.Tensor], n_units: int, group=None) -> torch.Tensor:
    """Generic form for the other generators (DETR (image, query) pairs, LXMERT / VisualBERT questions, ViT images:
    SURVEY.md §8e "independent by construction"): rank r calls ``unit_fn(lo, hi)`` -> maps ``[hi-lo, ...]`` for its
    contiguous slice of ``n_units`` and every rank receives all ``[n_units, ...]`` maps through the one all-gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(n_units, rank, world)
    if hi > lo:
        local = unit_fn(lo, hi)
    else:
        local = unit_fn(0, 1)[:0]
    if world == 1:
        return local
    return all_gather_maps(local, n_units, group)
