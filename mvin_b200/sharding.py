"""Host-side logic of the multi-GPU layouts -- pure torch, no CUDA needed, so it is covered by
world_size-2 `gloo` tests on CPU.

  * entity-table row sharding: entity e lives in shard e % G at local row e // G (include/mvin_b200.h:
    mvin_bind_entity_shards); every shard is padded to ceil(n_entity / G) rows.
  * data-parallel batch split: pairs are independent units; rank r takes the r-th contiguous slice.
  * replicated-gradient exchange: ONE flat SUM all-reduce of all small gradient tensors (+ the 4 loss scalars).
    With the loss scaling of mvin_set_batch_scale (base loss / global batch, dense L2 terms / world size) the
    summed result equals the single-device result on the concatenated batch.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def shard_rows(n_entity: int, n_shards: int) -> int:
    return (n_entity + n_shards - 1) // n_shards


def owner_of(e, n_shards: int):
    return e % n_shards


def local_row(e, n_shards: int):
    return e // n_shards


def scatter_table(full: torch.Tensor, n_shards: int, shard: int) -> torch.Tensor:
    """Rows of `full` [n_entity, d] owned by `shard`, zero-padded to shard_rows(n_entity, n_shards)."""
    part = full[shard::n_shards]
    out = torch.zeros((shard_rows(full.shape[0], n_shards),) + tuple(full.shape[1:]), dtype=full.dtype)
    out[:part.shape[0]] = part
    return out


def gather_table(parts: Sequence[torch.Tensor], n_entity: int) -> torch.Tensor:
    """Inverse of scatter_table over all shards."""
    G = len(parts)
    full = torch.zeros((parts[0].shape[0] * G,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype)
    for g, p in enumerate(parts):
        full[g::G] = p.cpu()
    return full[:n_entity]


def split_batch(n: int, rank: int, world: int) -> slice:
    """Contiguous slice of a global batch of n pairs owned by `rank` (n must divide evenly: the reference's batch
    size is static, model.py:251)."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by the world size {world}")
    per = n // world
    return slice(rank * per, (rank + 1) * per)


def allreduce_flat(tensors: List[torch.Tensor], group=None, extra: Optional[torch.Tensor] = None):
    """SUM all-reduce of `tensors` (in place) and of the optional 1-D `extra` (returned) with one collective."""
    import torch.distributed as dist
    dev = tensors[0].device if tensors else extra.device
    pieces = [t.reshape(-1) for t in tensors]
    if extra is not None:
        pieces.append(extra.to(dev, dtype=tensors[0].dtype if tensors else extra.dtype).reshape(-1))
    flat = torch.cat(pieces)
    dist.all_reduce(flat, group=group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].reshape(t.shape))
        off += t.numel()
    return flat[off:] if extra is not None else None
