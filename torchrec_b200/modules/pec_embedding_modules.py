"""Permissioned / position-encoded embedding collection ("PEC", reference torchrec/modules/pec_embedding_modules.py:26):
an EmbeddingCollection whose lookup is preceded by a per-feature id transformation hook and that returns embeddings
together with the remapped ids. Used where the id space is re-encoded per request (e.g. positional buckets)."""
from typing import Callable, Dict, Optional

import torch
from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .embedding_modules import EmbeddingCollection


class PECEmbeddingCollection(nn.Module):
    def __init__(self, embedding_collection: EmbeddingCollection, encoders: Optional[Dict[str, Callable[[torch.Tensor, torch.Tensor], torch.Tensor]]] = None,
                 return_encoded_features: bool = False) -> None:
        """encoders: feature -> fn(values, position_in_bag) -> new ids (must stay inside the table)."""
        super().__init__()
        self._embedding_collection = embedding_collection
        self._encoders = encoders or {}
        self._return_encoded = return_encoded_features

    def encode(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        if not self._encoders:
            return features
        from ..ops import jagged as J

        pos = J.offsets_range(features.offsets()[:-1].long(), features.values().numel())
        lpk = features.length_per_key()
        vals, poss = list(torch.split(features.values(), lpk)), torch.split(pos, lpk)
        for i, k in enumerate(features.keys()):
            if k in self._encoders:
                vals[i] = self._encoders[k](vals[i], poss[i]).to(vals[i].dtype)
        return KeyedJaggedTensor(keys=features.keys(), values=torch.cat(vals), lengths=features.lengths(), offsets=features.offsets(), weights=features.weights_or_none(),
                                 stride=features.stride(), length_per_key=lpk)

    def forward(self, features: KeyedJaggedTensor):
        enc = self.encode(features)
        out: Dict[str, JaggedTensor] = self._embedding_collection(enc)
        return (out, enc) if self._return_encoded else out
