"""Table-wise sharding of sequence (un-pooled) embeddings.

Reference: ``torchrec/distributed/sharding/tw_sequence_sharding.py`` - ``TwSequenceEmbeddingDist`` :50-113, ``TwSequenceEmbeddingSharding`` :116-161, inference
variants :164-289. Output: the sequence all-to-all sends every looked-up row back to the rank that owns its sample (the input all-to-all run backwards).
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
from .sequence_sharding import InferSequenceShardingContext, SequenceShardingContext
from .tw_sharding import BaseTwEmbeddingSharding, InferTwSparseFeaturesDist, TwSparseFeaturesDist


class TwSequenceEmbeddingDist(BaseEmbeddingDist[SequenceShardingContext, torch.Tensor, torch.Tensor]):
    def __init__(self, pg: dist.ProcessGroup, features_per_rank: List[int], device: Optional[torch.device] = None,
                 qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__()
        self._dist = SequenceEmbeddingsAllToAll(pg, features_per_rank, device, (qcomm_codecs_registry or {}).get(CommOp.SEQUENCE_EMBEDDINGS_ALL_TO_ALL.name))

    def forward(self, local_embs: torch.Tensor, sharding_ctx: Optional[SequenceShardingContext] = None) -> Awaitable[torch.Tensor]:
        assert sharding_ctx is not None
        return self._dist(local_embs, lengths=sharding_ctx.lengths_after_input_dist, input_splits=sharding_ctx.input_splits, output_splits=sharding_ctx.output_splits,
                          unbucketize_permute_tensor=None, batch_size_per_rank=sharding_ctx.batch_size_per_rank or None,
                          sparse_features_recat=sharding_ctx.sparse_features_recat)


class TwSequenceEmbeddingSharding(BaseTwEmbeddingSharding[SequenceShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[KeyedJaggedTensor]:
        assert self._pg is not None
        return TwSparseFeaturesDist(self._pg, self.features_per_rank())

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        assert feature_processor is None
        return self._sequence_lookup(device, fused_params)

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[SequenceShardingContext, torch.Tensor, torch.Tensor]:
        assert self._pg is not None
        return TwSequenceEmbeddingDist(self._pg, self.features_per_rank(), device if device is not None else self._device, self.qcomm_codecs_registry)


class InferTwSequenceEmbeddingDist(BaseEmbeddingDist[InferSequenceShardingContext, List[torch.Tensor], List[torch.Tensor]]):
    """Copy every device's rows to one device (they stay a list: features of different devices are different JaggedTensors)."""

    def __init__(self, device: torch.device, world_size: int) -> None:
        super().__init__()
        self._dist = SeqEmbeddingsAllToOne(device, world_size)

    def forward(self, local_embs: List[torch.Tensor], sharding_ctx: Optional[InferSequenceShardingContext] = None) -> List[torch.Tensor]:
        return self._dist(local_embs)


class InferTwSequenceEmbeddingSharding(BaseTwEmbeddingSharding[InferSequenceShardingContext, InputDistOutputs, List[torch.Tensor], List[torch.Tensor]]):
    def _copy_weights(self) -> None:
        self._init_rows = {}

    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[InputDistOutputs]:
        return InferTwSparseFeaturesDist(self.features_per_rank(), self._world_size, device)

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return InferGroupedEmbeddingsLookup(self._grouped_embedding_configs_per_rank, self._world_size, fused_params, device,
                                            device_type_from_sharding_infos=(device.type if device is not None else self._device.type))

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[InferSequenceShardingContext, List[torch.Tensor], List[torch.Tensor]]:
        return InferTwSequenceEmbeddingDist(device if device is not None else self._device, self._world_size)
