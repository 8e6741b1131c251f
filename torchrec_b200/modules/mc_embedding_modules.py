"""Managed-collision embedding collections (reference torchrec/modules/mc_embedding_modules.py:135-230):
remap ids through a ManagedCollisionCollection, look up, and reset the rows of evicted slots."""
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from .mc_modules import ManagedCollisionCollection


def evict(evictions: Dict[str, Optional[torch.Tensor]], ebc: Union[EmbeddingBagCollection, EmbeddingCollection]) -> None:
    """Re-initialise the embedding rows whose ids were evicted."""
    return


class BaseManagedCollisionEmbeddingCollection(nn.Module):
    def __init__(self, embedding_module: Union[EmbeddingBagCollection, EmbeddingCollection], managed_collision_collection: ManagedCollisionCollection,
                 return_remapped_features: bool = False, allow_in_place_embed_weight_update: bool = False) -> None:
        super().__init__()
        # rows of evicted ids are re-initialised by writing into the embedding weights; with autograd watching them that needs the version
        # counter bypass (``weight.data``) - allowed explicitly, as in the reference
        self._allow_in_place_embed_weight_update = allow_in_place_embed_weight_update
        self._managed_collision_collection = managed_collision_collection
        self._return_remapped_features = return_remapped_features
        self._embedding_module = embedding_module
        if isinstance(embedding_module, EmbeddingBagCollection):
            assert embedding_module.embedding_bag_configs() == managed_collision_collection.embedding_configs(), \
                "Embedding Bag Collection and Managed Collision Collection must contain the Embedding Configs"
        else:
            assert embedding_module.embedding_configs() == managed_collision_collection.embedding_configs(), \
                "Embedding Collection and Managed Collision Collection must contain the Embedding Configs"

    @torch.no_grad()
    def _reset_evicted(self) -> None:
        tables = self._embedding_module.embedding_bags if isinstance(self._embedding_module, EmbeddingBagCollection) else self._embedding_module.embeddings
        cfgs = {c.name: c for c in self._managed_collision_collection.embedding_configs()}
        for table, idx in self._managed_collision_collection.evict().items():
            if idx is None or idx.numel() == 0:
                continue
            w = tables[table].weight
            c = cfgs[table]
            w[idx.long().to(w.device)] = torch.empty(idx.numel(), w.shape[1], device=w.device, dtype=w.dtype).uniform_(c.get_weight_init_min(), c.get_weight_init_max())

    def forward(self, features: KeyedJaggedTensor):
        remapped = self._managed_collision_collection(features)
        if self.training:
            self._reset_evicted()
        out = self._embedding_module(remapped)
        return out, (remapped if self._return_remapped_features else None)


class ManagedCollisionEmbeddingCollection(BaseManagedCollisionEmbeddingCollection):
    def __init__(self, embedding_collection: EmbeddingCollection, managed_collision_collection: ManagedCollisionCollection, return_remapped_features: bool = False,
                 allow_in_place_embed_weight_update: bool = False) -> None:
        super().__init__(embedding_collection, managed_collision_collection, return_remapped_features, allow_in_place_embed_weight_update)

    @property
    def _embedding_collection(self) -> EmbeddingCollection:
        return self._embedding_module  # type: ignore[return-value]


class ManagedCollisionEmbeddingBagCollection(BaseManagedCollisionEmbeddingCollection):
    def __init__(self, embedding_bag_collection: EmbeddingBagCollection, managed_collision_collection: ManagedCollisionCollection, return_remapped_features: bool = False,
                 allow_in_place_embed_weight_update: bool = False) -> None:
        super().__init__(embedding_bag_collection, managed_collision_collection, return_remapped_features, allow_in_place_embed_weight_update)

    @property
    def _embedding_bag_collection(self) -> EmbeddingBagCollection:
        return self._embedding_module  # type: ignore[return-value]
