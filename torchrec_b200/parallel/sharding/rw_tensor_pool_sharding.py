"""Row-wise sharding of a TensorPool: rows ``[r * block, (r + 1) * block)`` of the pool live on rank ``r``.

Reference: ``torchrec/distributed/sharding/rw_tensor_pool_sharding.py`` - ``RwTensorPoolValuesDist`` :31-102, ``TensorPoolRwSharding`` :105-191,
``InferRwTensorPoolOutputDist`` :194-256, ``InferRwTensorPoolSharding`` :259-300. The sharded module (``parallel/object_pool.py: ShardedTensorPool``) uses the same
routing through its ``_Router``; these classes are the decomposed, reference-shaped form.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..dist_data import SeqEmbeddingsAllToOne
from ..tensor_sharding import InferObjectPoolSharding, ObjectPoolSharding, TensorPoolRwShardingContext
from ..types import Awaitable, NoWait, ShardingEnv
from .rw_pool_sharding import InferRwObjectPoolInputDist, RwObjectPoolIDsDist


class RwTensorPoolValuesDist(nn.Module):
    """Moves one row per id. ``is_update``: rows follow the ids TO the owners; otherwise looked-up rows travel BACK and are put in request order."""

    def __init__(self, pg: dist.ProcessGroup, is_update: bool) -> None:
        super().__init__()
        self._pg = pg
        self._is_update = is_update

    def forward(self, ctx: TensorPoolRwShardingContext, values: torch.Tensor) -> Awaitable[torch.Tensor]:
        if self._is_update:
            assert ctx.order is not None
            out = torch.empty(sum(ctx.recv_counts), *values.shape[1:], dtype=values.dtype, device=values.device)
            dist.all_to_all_single(out, values[ctx.order].contiguous(), output_split_sizes=ctx.recv_counts, input_split_sizes=ctx.send_counts, group=self._pg)
            return NoWait(out)
        back = torch.empty(sum(ctx.send_counts), *values.shape[1:], dtype=values.dtype, device=values.device)
        dist.all_to_all_single(back, values.contiguous(), output_split_sizes=ctx.send_counts, input_split_sizes=ctx.recv_counts, group=self._pg)
        assert ctx.unbucketize_permute is not None
        return NoWait(back[ctx.unbucketize_permute])


class TensorPoolRwSharding(ObjectPoolSharding):
    def __init__(self, pool_size: int, dim: int, env: ShardingEnv, device: torch.device) -> None:
        self._env = env
        self._pg = env.process_group
        self._world_size = env.world_size
        self._rank = env.rank
        self._device = device
        self._pool_size = pool_size
        self._dim = dim
        self._block_size = (pool_size + self._world_size - 1) // self._world_size
        self.local_pool_size = max(0, min(self._block_size, pool_size - self._rank * self._block_size))
        self._block_size_t = torch.tensor([self._block_size], device=device, dtype=torch.long)

    def create_update_ids_dist(self) -> RwObjectPoolIDsDist:
        return RwObjectPoolIDsDist(self._pg, is_update=True)

    def create_update_values_dist(self) -> RwTensorPoolValuesDist:
        return RwTensorPoolValuesDist(self._pg, is_update=True)

    def create_lookup_ids_dist(self) -> RwObjectPoolIDsDist:
        return RwObjectPoolIDsDist(self._pg, is_update=False)

    def create_lookup_values_dist(self) -> RwTensorPoolValuesDist:
        return RwTensorPoolValuesDist(self._pg, is_update=False)

    def get_sharded_states_to_register(self, lookup: nn.Module) -> Iterator[Tuple[str, torch.Tensor]]:
        yield from lookup.states_to_register()

    def create_context(self) -> TensorPoolRwShardingContext:
        return TensorPoolRwShardingContext(block_size=self._block_size_t)


class InferRwTensorPoolOutputDist(nn.Module):
    """Gather the per-device rows on one device and restore request order."""

    def __init__(self, env: ShardingEnv, device: torch.device) -> None:
        super().__init__()
        self._dist = SeqEmbeddingsAllToOne(device, env.world_size)

    def forward(self, lookups: List[torch.Tensor], unbucketize_permute: torch.Tensor) -> torch.Tensor:
        rows = torch.cat(self._dist(lookups), dim=0)
        return rows[unbucketize_permute.to(rows.device)]


class InferRwTensorPoolSharding(InferObjectPoolSharding):
    def create_lookup_ids_dist(self) -> InferRwObjectPoolInputDist:
        return InferRwObjectPoolInputDist(self._env, self._device, self._block_size_t)

    def create_lookup_values_dist(self) -> InferRwTensorPoolOutputDist:
        return InferRwTensorPoolOutputDist(self._env, self._device)
