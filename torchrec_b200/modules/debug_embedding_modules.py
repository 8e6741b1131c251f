"""Debug wrappers that validate / log the inputs and outputs of embedding collections
(reference torchrec/modules/debug_embedding_modules.py:46,133)."""
import logging
from typing import Dict, List

import torch
from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection

logger = logging.getLogger(__name__)


def _check(features: KeyedJaggedTensor, hash_sizes: Dict[str, int]) -> None:
    jt = features.to_dict()
    for k, f in jt.items():
        if k in hash_sizes and f.values().numel():
            mx, mn = int(f.values().max()), int(f.values().min())
            if mn < 0 or mx >= hash_sizes[k]:
                raise ValueError(f"feature {k}: ids in [{mn}, {mx}] outside of the table range [0, {hash_sizes[k]})")


class DebugEmbeddingBagCollection(nn.Module):
    def __init__(self, ebc: EmbeddingBagCollection) -> None:
        super().__init__()
        self._ebc = ebc
        self._hash_sizes = {f: c.num_embeddings for c in ebc.embedding_bag_configs() for f in c.feature_names}

    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        _check(features, self._hash_sizes)
        out = self._ebc(features)
        if not torch.isfinite(out.values()).all():
            raise ValueError("non-finite pooled embeddings")
        logger.debug(f"EBC out: keys={out.keys()} shape={tuple(out.values().shape)} mean={float(out.values().mean()):.5f}")
        return out


class DebugEmbeddingCollection(nn.Module):
    def __init__(self, ec: EmbeddingCollection) -> None:
        super().__init__()
        self._ec = ec
        self._hash_sizes = {f: c.num_embeddings for c in ec.embedding_configs() for f in c.feature_names}

    def forward(self, features: KeyedJaggedTensor) -> Dict[str, JaggedTensor]:
        _check(features, self._hash_sizes)
        out = self._ec(features)
        for k, v in out.items():
            if not torch.isfinite(v.values()).all():
                raise ValueError(f"non-finite embeddings for {k}")
        return out
