"""Position weights for all features of a group in one pass (reference ``torchrec/distributed/grouped_position_weighted.py:22``).

One flat parameter holds every feature's ``max_feature_length`` weights; ``forward`` gathers ``weight[feature_offset + position]`` for every id with a
single index (``offsets_range`` gives positions inside each bag) and returns the KJT with those weights attached."""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Tuple

import torch
from torch import nn

from ..modules.feature_processor import BaseGroupedFeatureProcessor
from ..ops import jagged as J
from ..sparse.jagged_tensor import KeyedJaggedTensor


class GroupedPositionWeightedModule(BaseGroupedFeatureProcessor):
    def __init__(self, max_feature_lengths: Dict[str, int], device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.max_feature_lengths = max_feature_lengths
        for length in max_feature_lengths.values():
            if length <= 0:
                raise ValueError("max_feature_length must be positive")
        self.position_weights = nn.ParameterDict()
        for key, length in max_feature_lengths.items():
            self.position_weights[key] = nn.Parameter(torch.ones(length, device=device))
        self.register_buffer("_dummy_weights", torch.tensor(max(max_feature_lengths.values()), device=device).fill_(1.0), persistent=False)

    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        if features.weights_or_none() is None:
            pos = J.offsets_range(features.offsets().long(), features.values().numel())
        else:  # row-wise input dist already replaced the weights by positions (bucketize_pos)
            pos = features.weights().long()
        lpk = features.length_per_key()
        weights_list, o = [], 0
        for key, n in zip(features.keys(), lpk):
            if key in self.max_feature_lengths:
                w = self.position_weights[key]
                weights_list.append(w[pos[o : o + n].clamp(max=w.numel() - 1)])
            else:
                weights_list.append(self._dummy_weights.expand(n))
            o += n
        weights = torch.cat(weights_list) if weights_list else features.values().new_empty(0, dtype=torch.float32)
        return KeyedJaggedTensor(keys=features.keys(), values=features.values(), weights=weights, lengths=features.lengths(), offsets=features.offsets(),
                                 stride=features.stride(), length_per_key=lpk)

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
        for name, p in self.position_weights.items():
            yield (f"{prefix}.position_weights.{name}" if prefix else f"position_weights.{name}"), p

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, torch.Tensor]]:
        yield from ()

    def state_dict(self, destination=None, prefix: str = "", keep_vars: bool = False):  # type: ignore[override]
        if destination is None:
            destination = {}
        for name, p in self.position_weights.items():
            destination[f"{prefix}position_weights.{name}"] = p if keep_vars else p.detach()
        return destination
