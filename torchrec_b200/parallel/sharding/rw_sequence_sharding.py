"""Row-wise sharding of sequence embeddings.

Reference: ``torchrec/distributed/sharding/rw_sequence_sharding.py`` - ``RwSequenceEmbeddingDist`` :57-118, ``RwSequenceEmbeddingSharding`` :121-202, inference
variants :205-344. Ids are bucketized by row owner before the all-to-all; after the output all-to-all the rows are in bucket order, and
``unbucketize_permute_tensor`` (kept in the sharding context) puts them back in the order of the original ids.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..dist_data import SeqEmbeddingsAllToOne, SequenceEmbeddingsAllToAll
from ..embedding_lookup import InferGroupedEmbeddingsLookup
from ..embedding_sharding import BaseEmbeddingDist, BaseSparseFeaturesDist
from ..embedding_types import BaseEmbeddingLookup, InputDistOutputs
from ..types import Awaitable, CommOp, QuantizedCommCodecs
from .rw_sharding import BaseRwEmbeddingSharding, InferRwSparseFeaturesDist, RwSparseFeaturesDist
from .sequence_sharding import InferSequenceShardingContext, SequenceShardingContext


class RwSequenceEmbeddingDist(BaseEmbeddingDist[SequenceShardingContext, torch.Tensor, torch.Tensor]):
    def __init__(self, pg: dist.ProcessGroup, num_features: int, device: Optional[torch.device] = None,
                 qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__()
        self._dist = SequenceEmbeddingsAllToAll(pg, [num_features] * pg.size(), device, (qcomm_codecs_registry or {}).get(CommOp.SEQUENCE_EMBEDDINGS_ALL_TO_ALL.name))

    def forward(self, local_embs: torch.Tensor, sharding_ctx: Optional[SequenceShardingContext] = None) -> Awaitable[torch.Tensor]:
        assert sharding_ctx is not None
        return self._dist(local_embs, lengths=sharding_ctx.lengths_after_input_dist, input_splits=sharding_ctx.input_splits, output_splits=sharding_ctx.output_splits,
                          unbucketize_permute_tensor=sharding_ctx.unbucketize_permute_tensor, batch_size_per_rank=sharding_ctx.batch_size_per_rank or None,
                          sparse_features_recat=sharding_ctx.sparse_features_recat)


class RwSequenceEmbeddingSharding(BaseRwEmbeddingSharding[SequenceShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[KeyedJaggedTensor]:
        assert self._pg is not None
        return RwSparseFeaturesDist(self._pg, self._get_num_features(), self._get_feature_hash_sizes(), self._get_feature_total_num_buckets(),
                                    device if device is not None else self._device, is_sequence=True, has_feature_processor=False, need_pos=False,
                                    embedding_shard_metadata=self._row_boundaries())

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return self._sequence_lookup(device, fused_params)

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[SequenceShardingContext, torch.Tensor, torch.Tensor]:
        assert self._pg is not None
        return RwSequenceEmbeddingDist(self._pg, self._get_num_features(), device if device is not None else self._device, self.qcomm_codecs_registry)

    def create_update(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None, feature_processor: Optional[nn.Module] = None) -> nn.Module:
        from ..embedding_lookup import GroupedEmbeddingsUpdate

        if not hasattr(self, "_update_lookup"):
            self._update_lookup = self._sequence_lookup(device, fused_params)
        return GroupedEmbeddingsUpdate(self._update_lookup, self._pg, device)

    def create_write_dist(self, device: Optional[torch.device] = None):
        from .rw_sharding import RwSparseFeaturesWriteDist

        assert self._pg is not None
        return RwSparseFeaturesWriteDist(self._pg, self._get_num_features(), self._get_feature_hash_sizes(), device if device is not None else self._device,
                                         embedding_shard_metadata=self._row_boundaries())


class InferRwSequenceEmbeddingDist(BaseEmbeddingDist[InferSequenceShardingContext, List[torch.Tensor], List[torch.Tensor]]):
    def __init__(self, device: torch.device, world_size: int, device_type_from_sharding_infos: Optional[str] = None) -> None:
        super().__init__()
        self._dist = SeqEmbeddingsAllToOne(device, world_size)

    def forward(self, local_embs: List[torch.Tensor], sharding_ctx: Optional[InferSequenceShardingContext] = None) -> List[torch.Tensor]:
        return self._dist(local_embs)


class InferRwSequenceEmbeddingSharding(BaseRwEmbeddingSharding[InferSequenceShardingContext, InputDistOutputs, List[torch.Tensor], List[torch.Tensor]]):
    def _copy_weights(self) -> None:
        self._init_rows = {}

    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[InputDistOutputs]:
        return InferRwSparseFeaturesDist(self._world_size, self._get_num_features(), self._get_feature_hash_sizes(), self._get_feature_total_num_buckets(),
                                         device if device is not None else self._device, is_sequence=True, embedding_shard_metadata=self._row_boundaries())

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return InferGroupedEmbeddingsLookup(self._grouped_embedding_configs_per_rank, self._world_size, fused_params, device,
                                            device_type_from_sharding_infos=(device.type if device is not None else self._device.type))

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[InferSequenceShardingContext, List[torch.Tensor], List[torch.Tensor]]:
        return InferRwSequenceEmbeddingDist(device if device is not None else self._device, self._world_size)
