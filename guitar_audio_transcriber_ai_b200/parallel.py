"""Clip sharding across the GPUs of one box and the label all-gather.

After slicing every clip is independent, so rank r simply owns a contiguous block of ceil(N/G) clips and
runs the whole pipeline on it; weights are replicated.  The only collective in the system is one
all-gather of a fixed-width record per clip (label index, confidence, slice start, slice end), padded so
every rank contributes the same count.  ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) does it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ceil(N/G) items for ``rank`` (the last ranks may be short or empty)."""
    per = -(-n_items // world_size) if n_items > 0 else 0
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def pack_records(indices: torch.Tensor, conf: torch.Tensor, table: torch.Tensor | None = None) -> torch.Tensor:
    """[n, 4] int64 records: label index, float32 confidence bits, start sample, end sample."""
    n = indices.shape[0]
    rec = torch.zeros((n, 4), dtype=torch.int64, device=indices.device)
    rec[:, 0] = indices.to(torch.int64)
    rec[:, 1] = conf.to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    if table is not None and n:
        rec[:, 2] = table[:, -2].to(rec.device)
        rec[:, 3] = table[:, -1].to(rec.device)
    return rec


def unpack_records(rec: torch.Tensor):
    idx = rec[:, 0]
    conf = rec[:, 1].to(torch.int32).view(torch.float32)
    return idx, conf, rec[:, 2], rec[:, 3]


def all_gather_records(rec: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Gathers every rank's [n_r, 4] block (rank r holds shard_bounds(n_total, G, r)) into [n_total, 4]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rec
    world = dist.get_world_size(group)
    per = -(-n_total // world)
    padded = torch.zeros((per, 4), dtype=torch.int64, device=rec.device)
    padded[: rec.shape[0]] = rec
    out = torch.empty((world * per, 4), dtype=torch.int64, device=rec.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_total]
