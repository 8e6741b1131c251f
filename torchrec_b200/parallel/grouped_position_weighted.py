"""Position weights of a whole feature group with ONE gather.

Role of the reference's ``GroupedPositionWeightedModule`` (``torchrec/distributed/grouped_position_weighted.py:22``): attach to every id of a KJT the learnt
weight of its position inside its bag. Design here: all features' weight vectors live in one flat parameter; a per-feature base offset is looked up for
every id (``repeat_interleave`` over ``length_per_key``), the position comes from ``offsets_range`` (or from the KJT's weights when the row-wise input
dist already wrote positions there), and a single ``index_select`` of ``base + min(pos, len - 1)`` produces all weights - no per-feature python loop on
the hot path, one kernel regardless of the number of features. Features without position weights read a constant 1 stored at slot 0.
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Tuple

import torch
from torch import nn

from ..modules.feature_processor import BaseGroupedFeatureProcessor
from ..ops import jagged as J
from ..sparse.jagged_tensor import KeyedJaggedTensor


class GroupedPositionWeightedModule(BaseGroupedFeatureProcessor):
    def __init__(self, max_feature_lengths: Dict[str, int], device: Optional[torch.device] = None) -> None:
        super().__init__()
        bad = [k for k, v in max_feature_lengths.items() if v <= 0]
        if bad:
            raise ValueError(f"max_feature_length must be positive (features {bad})")
        self.max_feature_lengths = dict(max_feature_lengths)
        self._slots: Dict[str, Tuple[int, int]] = {}
        start = 1  # slot 0 = the constant weight of features that are not position weighted
        for key, n in self.max_feature_lengths.items():
            self._slots[key] = (start, n)
            start += n
        self.flat_weights = nn.Parameter(torch.ones(start, device=device))
        self._plan_keys: Optional[List[str]] = None

    # -- the (base, limit) tables of one key order are cached: the KJT layout of a sharded module never changes --
    def _plan(self, keys: List[str], device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._plan_keys != keys or self._base.device != device:
            base = [self._slots.get(k, (0, 1))[0] for k in keys]
            limit = [self._slots.get(k, (0, 1))[1] - 1 for k in keys]
            self._base = torch.tensor(base, dtype=torch.long, device=device)
            self._limit = torch.tensor(limit, dtype=torch.long, device=device)
            self._plan_keys = list(keys)
        return self._base, self._limit

    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        values = features.values()
        n = values.numel()
        base, limit = self._plan(list(features.keys()), values.device)
        per_key = torch.as_tensor(features.length_per_key(), dtype=torch.long, device=values.device)
        existing = features.weights_or_none()
        pos = existing.long() if existing is not None else J.offsets_range(features.offsets().long(), n)
        idx = torch.repeat_interleave(base, per_key, output_size=n) + torch.minimum(pos, torch.repeat_interleave(limit, per_key, output_size=n))
        with torch.no_grad():
            self.flat_weights.data[0] = 1.0  # the pass-through slot is not learnt
        weights = self.flat_weights.index_select(0, idx)
        return KeyedJaggedTensor(keys=features.keys(), values=values, weights=weights, lengths=features.lengths(), offsets=features.offsets(), stride=features.stride(),
                                 length_per_key=features.length_per_key())

    # -- per-feature views under the reference's names: ``position_weights.{feature}`` --
    @property
    def position_weights(self) -> Dict[str, torch.Tensor]:
        return {k: self.flat_weights[a : a + n] for k, (a, n) in self._slots.items()}

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
        yield (f"{prefix}.flat_weights" if prefix else "flat_weights"), self.flat_weights

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, torch.Tensor]]:
        yield from ()

    def state_dict(self, destination: Optional[Dict[str, Any]] = None, prefix: str = "", keep_vars: bool = False) -> Dict[str, Any]:  # type: ignore[override]
        destination = {} if destination is None else destination
        for k, view in self.position_weights.items():
            destination[f"{prefix}position_weights.{k}"] = view if keep_vars else view.detach()
        return destination

    def load_state_dict(self, state_dict: Dict[str, Any], strict: bool = True) -> Any:  # type: ignore[override]
        missing = []
        with torch.no_grad():
            for k, (a, n) in self._slots.items():
                src = state_dict.get(f"position_weights.{k}")
                if src is None:
                    missing.append(f"position_weights.{k}")
                else:
                    self.flat_weights[a : a + n].copy_(src)
        if strict and missing:
            raise RuntimeError(f"missing keys {missing}")
        return torch.nn.modules.module._IncompatibleKeys(missing, [])
