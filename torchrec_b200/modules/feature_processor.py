"""Legacy feature-processor API (reference torchrec/modules/feature_processor.py:122)."""
from typing import Dict, List, Optional

import torch
from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .feature_processor_ import PositionWeightedModule


class BaseFeatureProcessor(nn.Module):
    def forward(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        raise NotImplementedError


class BaseGroupedFeatureProcessor(nn.Module):
    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        raise NotImplementedError


class PositionWeightedProcessor(BaseGroupedFeatureProcessor):
    """Grouped position weighting over a KJT (features without an entry keep weight 1)."""

    def __init__(self, max_feature_lengths: Dict[str, int], device: Optional[torch.device] = None) -> None:
        super().__init__()
        from .feature_processor_ import PositionWeightedModuleCollection

        self._impl = PositionWeightedModuleCollection(max_feature_lengths, device)
        self.max_feature_lengths = max_feature_lengths

    @property
    def position_weights(self):
        return self._impl.position_weights

    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        return self._impl(features)
