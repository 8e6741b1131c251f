"""Data-parallel sequence embeddings: local lookup, nothing to distribute.

Reference: ``torchrec/distributed/sharding/dp_sequence_sharding.py`` - ``DpSequenceEmbeddingDist`` :30-58, ``DpSequenceEmbeddingSharding`` :61-93.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
from torch import nn

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..embedding_sharding import BaseEmbeddingDist, BaseSparseFeaturesDist
from ..embedding_types import BaseEmbeddingLookup
from ..types import Awaitable, NoWait
from .dp_sharding import BaseDpEmbeddingSharding, DpSparseFeaturesDist
from .sequence_sharding import SequenceShardingContext


class DpSequenceEmbeddingDist(BaseEmbeddingDist[SequenceShardingContext, torch.Tensor, torch.Tensor]):
    def forward(self, local_embs: torch.Tensor, sharding_ctx: Optional[SequenceShardingContext] = None) -> Awaitable[torch.Tensor]:
        return NoWait(local_embs)


class DpSequenceEmbeddingSharding(BaseDpEmbeddingSharding[SequenceShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[KeyedJaggedTensor]:
        return DpSparseFeaturesDist()

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        assert feature_processor is None
        return self._sequence_lookup(device, fused_params, pg=self._env.process_group)

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[SequenceShardingContext, torch.Tensor, torch.Tensor]:
        return DpSequenceEmbeddingDist()
