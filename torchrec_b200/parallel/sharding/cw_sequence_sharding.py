"""Column-wise sharding of sequence embeddings: each column block is looked up by its owner, the sample owner concatenates the blocks.

Reference: ``torchrec/distributed/sharding/cw_sequence_sharding.py`` - ``CwSequenceEmbeddingSharding`` :40-84, inference variants :87-160. The dists are the
table-wise ones; the sharded module stitches a feature's blocks (``uncombined_embedding_names`` / ``embedding_shard_metadata`` give their column order).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
from torch import nn

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..embedding_sharding import BaseEmbeddingDist, BaseSparseFeaturesDist
from ..embedding_types import BaseEmbeddingLookup, InputDistOutputs
from .cw_sharding import BaseCwEmbeddingSharding
from .sequence_sharding import InferSequenceShardingContext, SequenceShardingContext
from .tw_sequence_sharding import InferTwSequenceEmbeddingDist, TwSequenceEmbeddingDist
from .tw_sharding import InferTwSparseFeaturesDist, TwSparseFeaturesDist


class CwSequenceEmbeddingSharding(BaseCwEmbeddingSharding[SequenceShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
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


class InferCwSequenceEmbeddingDist(InferTwSequenceEmbeddingDist):
    """reference :132-160."""


class InferCwSequenceEmbeddingSharding(BaseCwEmbeddingSharding[InferSequenceShardingContext, InputDistOutputs, List[torch.Tensor], List[torch.Tensor]]):
    def _copy_weights(self) -> None:
        self._init_rows = {}

    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[InputDistOutputs]:
        return InferTwSparseFeaturesDist(self.features_per_rank(), self._world_size, device)

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        from ..embedding_lookup import InferGroupedEmbeddingsLookup

        return InferGroupedEmbeddingsLookup(self._grouped_embedding_configs_per_rank, self._world_size, fused_params, device,
                                            device_type_from_sharding_infos=(device.type if device is not None else self._device.type))

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[InferSequenceShardingContext, List[torch.Tensor], List[torch.Tensor]]:
        return InferCwSequenceEmbeddingDist(device if device is not None else self._device, self._world_size)
