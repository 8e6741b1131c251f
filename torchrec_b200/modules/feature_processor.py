"""Legacy feature-processor API (reference torchrec/modules/feature_processor.py:122)."""
from typing import Dict, List, Optional

import torch
from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor


class BaseFeatureProcessor(nn.Module):
    def forward(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        raise NotImplementedError


class BaseGroupedFeatureProcessor(nn.Module):
    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        raise NotImplementedError


def position_weighted_module_update_features(features: Dict[str, JaggedTensor], weighted_features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
    """The input dict with the processed features replaced (features without position weights pass through)."""
    features.update(weighted_features)
    return features


def offsets_to_range_traceble(offsets: torch.Tensor, values: torch.Tensor) -> torch.Tensor:
    """Position of every value inside its bag (``[0, 1, 2, 0, 0, 1, ...]``)."""
    from ..ops import jagged as J

    return J.offsets_range(offsets[:-1].long() if offsets.numel() > 0 else offsets.long(), values.numel())


class PositionWeightedModule(BaseFeatureProcessor):
    """Legacy form over a dict of jagged tensors: every feature named in ``max_feature_lengths`` gets the weight
    ``position_weights[feature][position in bag]`` (learned, initialised to 1); other features pass through unchanged. Positions beyond
    the maximum length take the last weight. (The per-feature / KJT forms are in ``feature_processor_.py``.)"""

    def __init__(self, max_feature_lengths: Dict[str, int], device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.max_feature_lengths = max_feature_lengths
        self.position_weights = nn.ParameterDict({key: nn.Parameter(torch.ones(length, device=device)) for key, length in max_feature_lengths.items()})

    def reset_parameters(self) -> None:
        with torch.no_grad():
            for p in self.position_weights.values():
                p.fill_(1.0)

    def forward(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        weighted: Dict[str, JaggedTensor] = {}
        for key, w in self.position_weights.items():
            jt = features[key]
            pos = offsets_to_range_traceble(jt.offsets(), jt.values()).clamp(max=w.numel() - 1)
            weighted[key] = JaggedTensor(values=jt.values(), lengths=jt.lengths(), offsets=jt.offsets(), weights=torch.gather(w, 0, pos))
        return position_weighted_module_update_features(features, weighted)


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
