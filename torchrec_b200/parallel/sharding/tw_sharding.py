"""Table-wise sharding: a whole table on one rank.

Reference: ``torchrec/distributed/sharding/tw_sharding.py`` - ``BaseTwEmbeddingSharding`` :66-274, ``TwSparseFeaturesDist`` :277-315,
``TwPooledEmbeddingDist`` :318-415, ``TwPooledEmbeddingSharding`` :418-463, inference variants :466-584.
Input: KJT all-to-all with ``features_per_rank`` (the module feeds features in ``feature_names()`` order). Output: pooled all-to-all
``[B_global, D_local] -> [B_local, sum_r D_r]``; its variable-batch form when the batch carries per-feature batch sizes.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..dist_data import (
    EmbeddingsAllToOne,
    KJTAllToAll,
    KJTOneToAll,
    PooledEmbeddingsAllToAll,
    VariableBatchPooledEmbeddingsAllToAll,
)
from ..embedding_lookup import GroupedPooledEmbeddingsLookup, InferGroupedPooledEmbeddingsLookup
from ..embedding_sharding import (
    BaseEmbeddingDist,
    BaseSparseFeaturesDist,
    C,
    EmbeddingShardingContext,
    EmbeddingShardingInfo,
    F,
    T,
    W,
)
from ..embedding_types import BaseEmbeddingLookup, InputDistOutputs, KJTList, ShardedEmbeddingTable
from ..types import Awaitable, CommOp, NoWait, NullShardingContext, QuantizedCommCodecs, ShardingEnv
from .common import BaseShardingCommon, make_shard_table, rank_of, shards_of_info


class BaseTwEmbeddingSharding(BaseShardingCommon[C, F, T, W]):
    def _shard(self, sharding_infos: List[EmbeddingShardingInfo]) -> List[List[ShardedEmbeddingTable]]:
        tables_per_rank: List[List[ShardedEmbeddingTable]] = [[] for _ in range(self._world_size)]
        for info in sharding_infos:
            shards = shards_of_info(info)
            ranks = info.param_sharding.ranks or [rank_of(shards[0].placement)]
            rows, cols = shards[0].shard_sizes
            tables_per_rank[ranks[0]].append(make_shard_table(info, shards[0], rows, cols, _global_md(info)))
        return tables_per_rank


def _global_md(info: EmbeddingShardingInfo) -> Any:
    """ShardedTensorMetadata of the table when the plan carries a spec torch can describe; else None (plain tensors in state dicts)."""
    try:
        from torch.distributed._shard.sharded_tensor import ShardedTensorMetadata, TensorProperties

        cfg = info.embedding_config
        p = info.param
        dtype = p.dtype if isinstance(p, torch.Tensor) else torch.float32
        return ShardedTensorMetadata(shards_metadata=list(info.param_sharding.sharding_spec.shards), size=torch.Size([cfg.num_embeddings, cfg.embedding_dim]),  # type: ignore[union-attr]
                                     tensor_properties=TensorProperties(dtype=dtype))
    except Exception:
        return None


class TwSparseFeaturesDist(BaseSparseFeaturesDist[KeyedJaggedTensor]):
    """Whole features to their owner: one KJT all-to-all with ``features_per_rank`` splits."""

    def __init__(self, pg: dist.ProcessGroup, features_per_rank: List[int]) -> None:
        super().__init__()
        self._dist = KJTAllToAll(pg=pg, splits=features_per_rank)

    def forward(self, sparse_features: KeyedJaggedTensor) -> Awaitable[Awaitable[KeyedJaggedTensor]]:
        return self._dist(sparse_features)


class TwPooledEmbeddingDist(BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]):
    def __init__(self, pg: dist.ProcessGroup, dim_sum_per_rank: List[int], emb_dim_per_rank_per_feature: List[List[int]], device: Optional[torch.device] = None,
                 callbacks: Optional[List[Any]] = None, qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__()
        self._pg = pg
        self._dim_sum_per_rank = dim_sum_per_rank
        self._device = device
        self._callbacks = callbacks
        self._codecs = (qcomm_codecs_registry or {}).get(CommOp.POOLED_EMBEDDINGS_ALL_TO_ALL.name)
        self._emb_dim_per_rank_per_feature = emb_dim_per_rank_per_feature
        self._dist = PooledEmbeddingsAllToAll(pg, dim_sum_per_rank, device, callbacks, self._codecs)
        self._variable_dist: Optional[VariableBatchPooledEmbeddingsAllToAll] = None

    def forward(self, local_embs: torch.Tensor, sharding_ctx: Optional[EmbeddingShardingContext] = None) -> Awaitable[torch.Tensor]:
        if self._dist is None:
            return NoWait(local_embs)
        if sharding_ctx is None:
            return self._dist(local_embs)
        if sharding_ctx.variable_batch_per_feature:
            if self._variable_dist is None:
                self._variable_dist = VariableBatchPooledEmbeddingsAllToAll(self._pg, self._emb_dim_per_rank_per_feature, self._device, self._callbacks, self._codecs)
            return self._variable_dist(local_embs, sharding_ctx.batch_size_per_rank_per_feature, sharding_ctx.batch_size_per_feature_pre_a2a)
        return self._dist(local_embs, batch_size_per_rank=sharding_ctx.batch_size_per_rank or None)


class TwPooledEmbeddingSharding(BaseTwEmbeddingSharding[EmbeddingShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[KeyedJaggedTensor]:
        assert self._pg is not None
        return TwSparseFeaturesDist(self._pg, self.features_per_rank())

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return self._pooled_lookup(device, fused_params, feature_processor)

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]:
        assert self._pg is not None
        return TwPooledEmbeddingDist(self._pg, self._dim_sum_per_rank(), self._emb_dim_per_rank_per_feature(), device if device is not None else self._device,
                                     qcomm_codecs_registry=self.qcomm_codecs_registry)


# ---- inference: one process drives every device -------------------------------------------------------------------------
class InferTwSparseFeaturesDist(BaseSparseFeaturesDist[InputDistOutputs]):
    """Split the KJT by ``features_per_rank`` and copy part ``r`` to device ``r`` (reference :466-502)."""

    def __init__(self, features_per_rank: List[int], world_size: int, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._dist = KJTOneToAll(features_per_rank, world_size, device)

    def forward(self, sparse_features: KeyedJaggedTensor) -> InputDistOutputs:
        return InputDistOutputs(features=KJTList(self._dist(sparse_features).wait()))


class InferTwPooledEmbeddingDist(BaseEmbeddingDist[NullShardingContext, List[torch.Tensor], torch.Tensor]):
    """Gather the devices' pooled outputs on one device and concatenate the columns (reference :505-540)."""

    def __init__(self, device: torch.device, world_size: int) -> None:
        super().__init__()
        self._dist = EmbeddingsAllToOne(device, world_size, 1)

    def forward(self, local_embs: List[torch.Tensor], sharding_ctx: Optional[NullShardingContext] = None) -> torch.Tensor:
        return self._dist(local_embs)


class InferTwEmbeddingSharding(BaseTwEmbeddingSharding[NullShardingContext, InputDistOutputs, List[torch.Tensor], torch.Tensor]):
    def _copy_weights(self) -> None:  # quantized weights are loaded through the state dict
        self._init_rows = {}

    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[InputDistOutputs]:
        return InferTwSparseFeaturesDist(self.features_per_rank(), self._world_size, device)

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return InferGroupedPooledEmbeddingsLookup(self._grouped_embedding_configs_per_rank, self._world_size, fused_params, device, feature_processor,
                                                  device_type_from_sharding_infos=(device.type if device is not None else self._device.type))

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[NullShardingContext, List[torch.Tensor], torch.Tensor]:
        return InferTwPooledEmbeddingDist(device if device is not None else self._device, self._world_size)
