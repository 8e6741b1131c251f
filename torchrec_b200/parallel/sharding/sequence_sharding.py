"""Contexts of the sequence (un-pooled) shardings: what the output all-to-all must remember from the input dist.

Reference: ``torchrec/distributed/sharding/sequence_sharding.py`` - ``SequenceShardingContext`` :21-54, ``InferSequenceShardingContext`` :57-96.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ...streamable import Multistreamable
from ..embedding_sharding import EmbeddingShardingContext
from ..embedding_types import KJTList


class SequenceShardingContext(EmbeddingShardingContext):
    """``features_before_input_dist``: the KJT in this sharding's feature order (builds the output JaggedTensors); ``input_splits`` / ``output_splits``:
    rows sent / received per rank by the OUTPUT all-to-all; ``sparse_features_recat``: key permutation of the input all-to-all;
    ``unbucketize_permute_tensor``: restores the original id order after row-wise bucketization; ``lengths_after_input_dist``: per (feature, sample)
    lengths of what this rank looked up."""

    def __init__(self, batch_size_per_rank: Optional[List[int]] = None, features_before_input_dist: Optional[KeyedJaggedTensor] = None,
                 input_splits: Optional[List[int]] = None, output_splits: Optional[List[int]] = None, sparse_features_recat: Optional[torch.Tensor] = None,
                 unbucketize_permute_tensor: Optional[torch.Tensor] = None, lengths_after_input_dist: Optional[torch.Tensor] = None) -> None:
        super().__init__(batch_size_per_rank)
        self.features_before_input_dist = features_before_input_dist
        self.input_splits: List[int] = input_splits if input_splits is not None else []
        self.output_splits: List[int] = output_splits if output_splits is not None else []
        self.sparse_features_recat = sparse_features_recat
        self.unbucketize_permute_tensor = unbucketize_permute_tensor
        self.lengths_after_input_dist = lengths_after_input_dist

    def record_stream(self, stream: torch.Stream) -> None:
        if self.features_before_input_dist is not None:
            self.features_before_input_dist.record_stream(stream)
        for t in (self.sparse_features_recat, self.unbucketize_permute_tensor, self.lengths_after_input_dist):
            if t is not None and t.is_cuda:
                t.record_stream(stream)


class InferSequenceShardingContext(Multistreamable):
    """Inference: the per-device KJTs + the bookkeeping to stitch the per-device rows back into request order."""

    def __init__(self, features: KJTList, features_before_input_dist: Optional[KeyedJaggedTensor] = None, unbucketize_permute_tensor: Optional[torch.Tensor] = None,
                 bucket_mapping_tensor: Optional[torch.Tensor] = None, bucketized_length: Optional[torch.Tensor] = None,
                 embedding_names_per_rank: Optional[List[List[str]]] = None) -> None:
        super().__init__()
        self.features = features
        self.features_before_input_dist = features_before_input_dist
        self.unbucketize_permute_tensor = unbucketize_permute_tensor
        self.bucket_mapping_tensor = bucket_mapping_tensor
        self.bucketized_length = bucketized_length
        self.embedding_names_per_rank = embedding_names_per_rank

    def record_stream(self, stream: torch.Stream) -> None:
        self.features.record_stream(stream)
        if self.features_before_input_dist is not None:
            self.features_before_input_dist.record_stream(stream)
        for t in (self.unbucketize_permute_tensor, self.bucket_mapping_tensor, self.bucketized_length):
            if t is not None and t.is_cuda:
                t.record_stream(stream)
