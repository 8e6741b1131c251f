"""Data-parallel sequence embeddings (role of ``torchrec/distributed/sharding/dp_sequence_sharding.py``): the table is replicated, so the three stages
collapse to "look up locally". The input and output stages are the no-op dists of ``dp_sharding.py``; only the lookup differs from the pooled flavour
(sequence kernel, rows instead of bags). Kept as its own class so a sharded ``EmbeddingCollection`` can treat every sharding type uniformly."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
from torch import nn

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..embedding_types import BaseEmbeddingLookup
from .dp_sharding import BaseDpEmbeddingSharding, DpPooledEmbeddingDist, DpSparseFeaturesDist
from .sequence_sharding import SequenceShardingContext


class DpSequenceEmbeddingDist(DpPooledEmbeddingDist):
    """Identity on ``[sum(lengths), D]`` rows (already-waited awaitable, like the pooled flavour)."""


class DpSequenceEmbeddingSharding(BaseDpEmbeddingSharding[SequenceShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        if feature_processor is not None:
            raise ValueError("sequence embeddings take no feature processor")
        return self._sequence_lookup(device, fused_params, pg=self._env.process_group)

    def create_input_dist(self, device: Optional[torch.device] = None) -> DpSparseFeaturesDist:
        return DpSparseFeaturesDist()

    def create_output_dist(self, device: Optional[torch.device] = None) -> DpSequenceEmbeddingDist:
        return DpSequenceEmbeddingDist()
