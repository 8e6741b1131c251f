"""Embedding towers: an embedding module + the interaction that consumes it, kept on the same ranks
(reference torchrec/modules/embedding_tower.py:39,86)."""
from typing import List, Optional, Tuple, Union

import torch
from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor
from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection


def tower_input_params(module: nn.Module) -> Tuple[bool, bool]:
    if isinstance(module, EmbeddingCollection):
        return True, False
    if isinstance(module, EmbeddingBagCollection):
        return not module.is_weighted(), module.is_weighted()
    return True, True


class EmbeddingTower(nn.Module):
    def __init__(self, embedding_module: nn.Module, interaction_module: nn.Module, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.embedding = embedding_module
        self.interaction = interaction_module

    def forward(self, *args, **kwargs) -> torch.Tensor:
        return self.interaction(self.embedding(*args, **kwargs))


class EmbeddingTowerCollection(nn.Module):
    def __init__(self, towers: List[EmbeddingTower], device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.towers = nn.ModuleList(towers)
        self._input_params: List[Tuple[bool, bool]] = []
        for tower in towers:
            self._input_params.append(tower_input_params(tower.embedding))

    def forward(self, features: Optional[KeyedJaggedTensor] = None, weighted_features: Optional[KeyedJaggedTensor] = None) -> torch.Tensor:
        tower_outputs = []
        for tower, input_params in zip(self.towers, self._input_params):
            has_kjt_param, has_wkjt_param = input_params
            if has_kjt_param and has_wkjt_param:
                assert features is not None and weighted_features is not None
                tower_outputs.append(tower(features, weighted_features))
            elif has_wkjt_param:
                assert weighted_features is not None
                tower_outputs.append(tower(weighted_features))
            else:
                assert features is not None
                tower_outputs.append(tower(features))
        return torch.cat(tower_outputs, dim=1)
