"""Feature processors: position-weighted features (reference torchrec/modules/feature_processor_.py:52-260)."""
from __future__ import annotations

import abc
from typing import Dict, List, Optional

import torch
from torch import nn

from ..ops import jagged as J
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from ..types import CopyMixIn


class FeatureProcessor(nn.Module):
    """Transforms the weights of one JaggedTensor feature."""

    @abc.abstractmethod
    def forward(self, features: JaggedTensor) -> JaggedTensor:
        ...


class PositionWeightedModule(FeatureProcessor):
    """weight of the k-th id of every bag = learned ``position_weight[k]``."""

    def __init__(self, max_feature_length: int, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.position_weight = nn.Parameter(torch.empty([max_feature_length], device=device))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        with torch.no_grad():
            self.position_weight.fill_(1.0)

    def forward(self, features: JaggedTensor) -> JaggedTensor:
        seq = J.offsets_range(features.offsets()[:-1].long(), features.values().numel())
        seq = seq.clamp(max=self.position_weight.numel() - 1)
        weighted = torch.gather(self.position_weight, dim=0, index=seq)
        return JaggedTensor(values=features.values(), weights=weighted, lengths=features.lengths(), offsets=features.offsets())


class FeatureProcessorsCollection(nn.Module):
    """Transforms the weights of a whole KJT."""

    @abc.abstractmethod
    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        ...


def get_weights_list(cat_seq: torch.Tensor, features: KeyedJaggedTensor, position_weights: Dict[str, nn.Parameter]) -> Optional[torch.Tensor]:
    weights_list = []
    seqs = torch.split(cat_seq, features.length_per_key())
    for key, seq in zip(features.keys(), seqs):
        if key in position_weights:
            weights_list.append(torch.gather(position_weights[key], dim=0, index=seq.clamp(max=position_weights[key].numel() - 1)))
        else:
            weights_list.append(torch.ones(seq.shape[0], device=features.values().device, dtype=torch.float32))
    return torch.cat(weights_list) if weights_list else features.weights_or_none()


def get_stride_per_key_per_rank(kjt: KeyedJaggedTensor) -> Optional[List[List[int]]]:
    """The per-key per-rank batch sizes of a variable-batch KJT, None for a fixed batch."""
    return kjt.stride_per_key_per_rank() if kjt.variable_stride_per_key() else None


class PositionWeightedModuleCollection(FeatureProcessorsCollection, CopyMixIn):
    """One position-weight vector per feature (``max_feature_lengths``: feature -> max length)."""

    def __init__(self, max_feature_lengths: Dict[str, int], device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.max_feature_lengths = max_feature_lengths
        for length in self.max_feature_lengths.values():
            if length <= 0:
                raise ValueError("max_feature_length must be positive")
        self.position_weights: nn.ParameterDict = nn.ParameterDict()
        self.position_weights_dict: Dict[str, nn.Parameter] = {}
        for key, length in max_feature_lengths.items():
            self.position_weights[key] = nn.Parameter(torch.empty([length], device=device))
            self.position_weights_dict[key] = self.position_weights[key]
        self.reset_parameters()

    def reset_parameters(self) -> None:
        with torch.no_grad():
            for key, _ in self.max_feature_lengths.items():
                self.position_weights[key].fill_(1.0)
                self.position_weights_dict[key] = self.position_weights[key]

    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        cat_seq = J.offsets_range(features.offsets()[:-1].long(), features.values().numel())
        return KeyedJaggedTensor(
            keys=features.keys(), values=features.values(), weights=get_weights_list(cat_seq, features, self.position_weights_dict),
            lengths=features.lengths(), offsets=features.offsets(), stride=features.stride(), length_per_key=features.length_per_key())

    def copy(self, device: torch.device) -> nn.Module:
        self.position_weights = self.position_weights.to(device=device)
        for key in self.position_weights.keys():
            self.position_weights_dict[key] = self.position_weights[key]
        return self

    def _apply(self, *args, **kwargs) -> nn.Module:
        super()._apply(*args, **kwargs)
        for k, param in self.position_weights.items():
            self.position_weights_dict[k] = param
        return self
