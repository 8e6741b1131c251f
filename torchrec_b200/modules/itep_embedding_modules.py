"""Embedding collections with in-training embedding pruning: ids are remapped through the ITEP module before the lookup (reference ``torchrec/modules/itep_embedding_modules.py:24,88``)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple
import torch
from torch import nn
from ..sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor
from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from .itep_modules import GenericITEPModule  # noqa: F401


class ITEPEmbeddingBagCollection(nn.Module):
    def __init__(self, embedding_bag_collection: EmbeddingBagCollection, itep_module: GenericITEPModule) -> None:
        super().__init__()
        self._embedding_bag_collection = embedding_bag_collection
        self._itep_module = itep_module
        if not itep_module.feature_to_table:
            itep_module.feature_to_table = {f: c.name for c in embedding_bag_collection.embedding_bag_configs() for f in c.feature_names}
        self.register_buffer("_iter", torch.tensor(0, dtype=torch.int64))

    def forward(self, features: KeyedJaggedTensor, force_insert: bool = False) -> KeyedTensor:
        before = self._itep_module.num_prunes
        features = self._itep_module(features, int(self._iter.item()))
        if self._itep_module.num_prunes != before:  # rows changed owner in this call: they start from a fresh embedding
            self._itep_module.reset_weight_momentum(self._embedding_bag_collection)
        out = self._embedding_bag_collection(features)
        if self.training:
            self._iter += 1
        return out


class ITEPEmbeddingCollection(nn.Module):
    def __init__(self, embedding_collection: EmbeddingCollection, itep_module: GenericITEPModule) -> None:
        super().__init__()
        self._embedding_collection = embedding_collection
        self._itep_module = itep_module
        if not itep_module.feature_to_table:
            itep_module.feature_to_table = {f: c.name for c in embedding_collection.embedding_configs() for f in c.feature_names}
        self.register_buffer("_iter", torch.tensor(0, dtype=torch.int64))

    def forward(self, features: KeyedJaggedTensor, force_insert: bool = False):
        before = self._itep_module.num_prunes
        features = self._itep_module(features, int(self._iter.item()))
        if self._itep_module.num_prunes != before:
            self._itep_module.reset_weight_momentum(self._embedding_collection)
        out = self._embedding_collection(features)
        if self.training:
            self._iter += 1
        return out
