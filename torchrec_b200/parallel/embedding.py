"""Placeholder — replaced below by the full ShardedEmbeddingCollection."""
from typing import Type

from ..modules.embedding_modules import EmbeddingCollection
from .embedding_types import BaseEmbeddingSharder


class EmbeddingCollectionSharder(BaseEmbeddingSharder[EmbeddingCollection]):
    @property
    def module_type(self) -> Type[EmbeddingCollection]:
        return EmbeddingCollection

    def shard(self, *a, **k):
        raise NotImplementedError
