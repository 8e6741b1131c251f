"""Ids dist of row-wise sharded object pools: id -> owner ``id // block_size``, ids sorted by owner, two all-to-alls (counts, ids).

Reference: ``torchrec/distributed/sharding/rw_pool_sharding.py`` - ``RwObjectPoolIDsDist`` :21-141, ``InferRwObjectPoolInputDist`` :157-230.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..tensor_sharding import ObjectPoolRwShardingContext
from ..types import Awaitable, NoWait


class RwObjectPoolIDsDist(nn.Module):
    def __init__(self, pg: dist.ProcessGroup, is_update: bool = True, bucketize_world_size: Optional[int] = None, bucketize_rank_offset: int = 0) -> None:
        super().__init__()
        self._pg = pg
        self._world_size = dist.get_world_size(pg)
        self._is_update = is_update
        self._buckets = bucketize_world_size or self._world_size
        self._rank_offset = bucketize_rank_offset

    def forward(self, ctx: ObjectPoolRwShardingContext, ids: torch.Tensor) -> Awaitable[Awaitable[torch.Tensor]]:
        assert ctx.block_size is not None
        ids = ids.long()
        block = int(ctx.block_size.reshape(-1)[0])
        owner = torch.div(ids, block, rounding_mode="floor").clamp(max=self._buckets - 1) + self._rank_offset
        order = torch.argsort(owner, stable=True)
        send = torch.bincount(owner, minlength=self._world_size)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self._pg)
        ctx.ids_before_input_dist = ids
        ctx.num_ids_each_rank_to_send, ctx.num_ids_each_rank_to_receive = send, recv
        ctx.send_counts, ctx.recv_counts = send.tolist(), recv.tolist()
        ctx.order = order
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel(), device=order.device)
        ctx.unbucketize_permute = inv
        out = torch.empty(sum(ctx.recv_counts), dtype=ids.dtype, device=ids.device)
        dist.all_to_all_single(out, ids[order].contiguous(), output_split_sizes=ctx.recv_counts, input_split_sizes=ctx.send_counts, group=self._pg)
        rank = dist.get_rank(self._pg)
        return NoWait(NoWait(out - (rank - self._rank_offset) * block))


def _get_bucketize_shape(ids: torch.Tensor, device: torch.device) -> torch.Tensor:
    return torch.tensor([ids.size(dim=0)], device=device, dtype=torch.long)


def _get_unbucketize_permute_index(unbucketize_permute: Optional[torch.Tensor]) -> torch.Tensor:
    assert unbucketize_permute is not None, "unbucketize permute must not be None"
    return unbucketize_permute.long()


class InferRwObjectPoolInputDist(nn.Module):
    """Single-process inference: split the ids by owner device and copy each piece there; returns (ids per device, permutation back to request order)."""

    def __init__(self, env, device: torch.device, block_size: torch.Tensor) -> None:
        super().__init__()
        self._world_size = env.world_size
        self._device = device
        self._block_size = block_size

    def forward(self, ids: torch.Tensor) -> Tuple[List[torch.Tensor], torch.Tensor]:
        ids = ids.long()
        block = int(self._block_size.reshape(-1)[0])
        owner = torch.div(ids, block, rounding_mode="floor").clamp(max=self._world_size - 1)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self._world_size).tolist()
        local = (ids - owner * block)[order]
        pieces = list(torch.split(local, counts))
        if self._device.type == "cuda":
            pieces = [p.to(torch.device("cuda", r), non_blocking=True) for r, p in enumerate(pieces)]
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel(), device=order.device)
        return pieces, inv
