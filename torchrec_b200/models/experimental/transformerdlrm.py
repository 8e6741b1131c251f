"""DLRM whose interaction layer is a transformer encoder over the feature embeddings (reference
torchrec/models/experimental/transformerdlrm.py:18-190; a benchmarking model: embeddings + attention in one step, not a quality claim)."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from ...modules.embedding_modules import EmbeddingBagCollection
from ..dlrm import DLRM, OverArch


class InteractionTransformerArch(nn.Module):
    """``[dense ; sparse_1 .. sparse_F]`` (``F + 1`` tokens of width D) -> ``nn.TransformerEncoder`` -> flattened ``[B, (F + 1) * D]``."""

    def __init__(self, num_sparse_features: int, embedding_dim: int, nhead: int = 8, ntransformer_layers: int = 4) -> None:
        super().__init__()
        self.F: int = num_sparse_features
        self.nhead = nhead
        self.ntransformer_layers = ntransformer_layers
        layer = nn.TransformerEncoderLayer(d_model=embedding_dim, nhead=nhead)
        self.interarch_TE = nn.TransformerEncoder(layer, num_layers=ntransformer_layers)

    def forward(self, dense_features: torch.Tensor, sparse_features: torch.Tensor) -> torch.Tensor:
        if self.F <= 0:
            return dense_features
        B = dense_features.shape[0]
        tokens = torch.cat((dense_features.unsqueeze(1), sparse_features.to(dense_features.dtype)), dim=1)  # [B, F + 1, D]
        # the encoder layers are sequence-first (batch_first=False, as in the reference): attention runs over dim 0. The reference
        # feeds [B, F + 1, D] as is, i.e. it attends ACROSS THE BATCH for every feature slot; kept for parity of the benchmark.
        out = self.interarch_TE(tokens)
        return out.reshape(B, -1)


class DLRM_Transformer(DLRM):
    def __init__(self, embedding_bag_collection: EmbeddingBagCollection, dense_in_features: int, dense_arch_layer_sizes: List[int], over_arch_layer_sizes: List[int],
                 nhead: int = 8, ntransformer_layers: int = 4, dense_device: Optional[torch.device] = None) -> None:
        super().__init__(embedding_bag_collection, dense_in_features, dense_arch_layer_sizes, over_arch_layer_sizes, dense_device)
        embedding_dim = embedding_bag_collection.embedding_bag_configs()[0].embedding_dim
        num_sparse_features = len(self.sparse_arch.sparse_feature_names)
        self.inter_arch = InteractionTransformerArch(num_sparse_features=num_sparse_features, embedding_dim=embedding_dim, nhead=nhead,
                                                     ntransformer_layers=ntransformer_layers)
        if dense_device is not None:
            self.inter_arch = self.inter_arch.to(dense_device)
        self.over_arch = OverArch(in_features=(num_sparse_features + 1) * embedding_dim, layer_sizes=over_arch_layer_sizes, device=dense_device)
