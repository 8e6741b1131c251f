"""Prioritized Embedding Communication (PEC) wrapper of an ``EmbeddingCollection`` (reference torchrec/modules/pec_embedding_modules.py:26-116).

PEC splits every batch's distributed ids into those that also occurred in the previous batch ("overlapped": their rows are being
updated by the previous step's backward and must be looked up after it) and the rest ("non-overlapped": their lookup can start while
the previous step is still running). The unsharded module only carries the configuration and delegates to the wrapped collection;
the logic lives in ``parallel/pec_embedding.py: ShardedPECEmbeddingCollection``."""
from enum import Enum
from typing import Dict, List

from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .embedding_configs import EmbeddingConfig
from .embedding_modules import EmbeddingCollection


class OverlappingCheckerType(Enum):
    """How overlap with the previous batch is detected. BOOLEAN: a boolean mask over the local rows of every shard."""

    BOOLEAN = "boolean"


class PECEmbeddingCollection(nn.Module):
    def __init__(self, embedding_collection: EmbeddingCollection, checker_type: OverlappingCheckerType = OverlappingCheckerType.BOOLEAN) -> None:
        super().__init__()
        self._embedding_collection = embedding_collection
        self._checker_type = OverlappingCheckerType(checker_type)

    @property
    def embedding_collection(self) -> EmbeddingCollection:
        return self._embedding_collection

    @property
    def checker_type(self) -> OverlappingCheckerType:
        return self._checker_type

    def forward(self, features: KeyedJaggedTensor) -> Dict[str, JaggedTensor]:
        return self._embedding_collection(features)

    def embedding_configs(self) -> List[EmbeddingConfig]:
        return self._embedding_collection.embedding_configs()

    def embedding_dim(self) -> int:
        return self._embedding_collection.embedding_dim()

    def need_indices(self) -> bool:
        return self._embedding_collection.need_indices()
