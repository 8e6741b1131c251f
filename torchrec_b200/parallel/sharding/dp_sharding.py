"""Data-parallel "sharding": every rank holds the whole table, no communication on the embedding path.

Reference: ``torchrec/distributed/sharding/dp_sharding.py`` - ``BaseDpEmbeddingSharding`` :41-133, ``DpSparseFeaturesDist`` :136-161,
``DpPooledEmbeddingDist`` :164-192, ``DpPooledEmbeddingSharding`` :195-231. The lookup uses the dense kernel (``OptimType.NONE``): its flat weight is an
autograd Parameter the sharded module wraps in DDP, so the gradient all-reduce is DDP's bucketed NCCL all-reduce (NVLS in-switch reduction on NVSwitch).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
from torch import nn

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..embedding_sharding import BaseEmbeddingDist, BaseSparseFeaturesDist, C, EmbeddingShardingContext, EmbeddingShardingInfo, F, T, W
from ..embedding_types import BaseEmbeddingLookup, EmbeddingComputeKernel, ShardedEmbeddingTable
from ..types import Awaitable, NoWait
from .common import BaseShardingCommon, make_shard_table


class BaseDpEmbeddingSharding(BaseShardingCommon[C, F, T, W]):
    def _shard(self, sharding_infos: List[EmbeddingShardingInfo]) -> List[List[ShardedEmbeddingTable]]:
        tables_per_rank: List[List[ShardedEmbeddingTable]] = [[] for _ in range(self._world_size)]
        for info in sharding_infos:
            cfg = info.embedding_config
            for rank in range(self._world_size):
                tables_per_rank[rank].append(make_shard_table(info, None, cfg.num_embeddings, cfg.embedding_dim))
        return tables_per_rank

    # every rank sees every feature: names are those of ONE rank
    def embedding_dims(self) -> List[int]:
        return [d for g in self._grouped_embedding_configs for d in g.embedding_dims()]

    def embedding_names(self) -> List[str]:
        return [n for g in self._grouped_embedding_configs for n in g.embedding_names()]

    def embedding_names_per_rank(self) -> List[List[str]]:
        raise NotImplementedError("data-parallel tables have no per-rank names")

    def embedding_shard_metadata(self) -> List[Optional[Any]]:
        return [None for g in self._grouped_embedding_configs for _ in g.embedding_names()]

    def feature_names(self) -> List[str]:
        return [f for g in self._grouped_embedding_configs for f in g.feature_names()]

    def embedding_tables(self) -> List[ShardedEmbeddingTable]:
        return [t for g in self._grouped_embedding_configs for t in g.embedding_tables]


class DpSparseFeaturesDist(BaseSparseFeaturesDist[KeyedJaggedTensor]):
    """No-op input dist (already-waited awaitables keep the module's three-stage call shape)."""

    def forward(self, sparse_features: KeyedJaggedTensor) -> Awaitable[Awaitable[KeyedJaggedTensor]]:
        return NoWait(NoWait(sparse_features))


class DpPooledEmbeddingDist(BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]):
    def forward(self, local_embs: torch.Tensor, sharding_ctx: Optional[EmbeddingShardingContext] = None) -> Awaitable[torch.Tensor]:
        return NoWait(local_embs)


class DpPooledEmbeddingSharding(BaseDpEmbeddingSharding[EmbeddingShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[KeyedJaggedTensor]:
        return DpSparseFeaturesDist()

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        # the replica group, not the sharding group, scales these gradients (DDP averages them): no comm-op gradient scaling
        return self._pooled_lookup(device, fused_params, feature_processor, pg=self._env.process_group, scale_weight_gradients=False)

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]:
        return DpPooledEmbeddingDist()
