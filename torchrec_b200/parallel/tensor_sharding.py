"""Sharding contexts and ABCs of the object pools (TensorPool / KeyedJaggedTensorPool).

Reference: ``torchrec/distributed/tensor_sharding.py`` - ``ObjectPoolShardingContext`` :22, ``RwShardingContext`` :43, ``ObjectPoolRwShardingContext`` :52,
``ObjectPoolReplicatedRwShardingContext`` :58, ``TensorPoolRwShardingContext`` :64, ``ObjectPoolSharding`` :73, ``InferObjectPoolSharding`` :99.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn

from ..streamable import Multistreamable
from .types import ShardingEnv


class ObjectPoolShardingContext(Multistreamable):
    """What the values / output dists need from the ids dist of the same call: the permutation that grouped ids by owner, the per-rank counts."""

    def __init__(self, ids_before_input_dist: Optional[torch.Tensor] = None, num_ids_each_rank_to_receive: Optional[torch.Tensor] = None,
                 num_ids_each_rank_to_send: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.ids_before_input_dist = ids_before_input_dist
        self.num_ids_each_rank_to_receive = num_ids_each_rank_to_receive
        self.num_ids_each_rank_to_send = num_ids_each_rank_to_send

    def record_stream(self, stream: torch.Stream) -> None:
        for t in (self.ids_before_input_dist, self.num_ids_each_rank_to_receive, self.num_ids_each_rank_to_send):
            if t is not None and t.is_cuda:
                t.record_stream(stream)


class RwShardingContext(Multistreamable):
    """Row-wise routing bookkeeping: ``unbucketize_permute`` restores the caller's id order from owner-sorted order."""

    def __init__(self, block_size: Optional[torch.Tensor] = None, unbucketize_permute: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.block_size = block_size
        self.unbucketize_permute = unbucketize_permute

    def record_stream(self, stream: torch.Stream) -> None:
        for t in (self.block_size, self.unbucketize_permute):
            if t is not None and t.is_cuda:
                t.record_stream(stream)


class ObjectPoolRwShardingContext(ObjectPoolShardingContext, RwShardingContext):
    def __init__(self, **kw) -> None:
        ObjectPoolShardingContext.__init__(self, kw.pop("ids_before_input_dist", None), kw.pop("num_ids_each_rank_to_receive", None), kw.pop("num_ids_each_rank_to_send", None))
        RwShardingContext.__init__(self, kw.pop("block_size", None), kw.pop("unbucketize_permute", None))
        # host copies of the counts (the collectives need python lists)
        self.send_counts: List[int] = []
        self.recv_counts: List[int] = []
        self.order: Optional[torch.Tensor] = None

    def record_stream(self, stream: torch.Stream) -> None:
        ObjectPoolShardingContext.record_stream(self, stream)
        RwShardingContext.record_stream(self, stream)


class ObjectPoolReplicatedRwShardingContext(ObjectPoolRwShardingContext):
    pass


class TensorPoolRwShardingContext(ObjectPoolRwShardingContext):
    pass


class ObjectPoolSharding(ABC):
    @abstractmethod
    def create_update_ids_dist(self) -> nn.Module:
        ...

    @abstractmethod
    def create_update_values_dist(self) -> nn.Module:
        ...

    @abstractmethod
    def create_lookup_ids_dist(self) -> nn.Module:
        ...

    @abstractmethod
    def create_lookup_values_dist(self) -> nn.Module:
        ...

    @abstractmethod
    def get_sharded_states_to_register(self, lookup: nn.Module):
        ...

    @abstractmethod
    def create_context(self) -> ObjectPoolShardingContext:
        ...


class InferObjectPoolSharding(ABC):
    def __init__(self, pool_size: int, env: ShardingEnv, device: torch.device) -> None:
        self._pool_size = pool_size
        self._env = env
        self._world_size = env.world_size
        self._device = device
        self._block_size = (pool_size + self._world_size - 1) // self._world_size
        self._last_block_size = pool_size - self._block_size * (self._world_size - 1)
        self.local_pool_size_per_rank = [self._block_size] * (self._world_size - 1) + [self._last_block_size]
        self._block_size_t = torch.tensor([self._block_size], device=device, dtype=torch.long)

    @abstractmethod
    def create_lookup_ids_dist(self) -> nn.Module:
        ...

    @abstractmethod
    def create_lookup_values_dist(self) -> nn.Module:
        ...
