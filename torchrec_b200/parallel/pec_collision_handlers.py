"""Overlap ("collision") detection between consecutive batches for PEC (reference torchrec/distributed/pec_collision_handlers.py).

The handler works on the OWNER side: on the distributed KJT a rank received for its lookup units. A row key is
``unit_row_base[unit] + local_id``; the boolean checker keeps one bool per local row, set by the previous batch's keys."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from ..modules.pec_embedding_modules import OverlappingCheckerType
from ..sparse.jagged_tensor import KeyedJaggedTensor


@dataclass
class CollisionResult:
    """``forward_overlap_mask[i]``: value i of THIS batch also occurred in the previous batch. ``backward_overlap_mask``: per value of
    the PREVIOUS batch, whether it occurs in this one (None on the first batch). ``remapped_feature_values``: this batch's row keys
    (what the next batch is compared against)."""

    forward_overlap_mask: torch.Tensor
    backward_overlap_mask: Optional[torch.Tensor]
    remapped_feature_values: torch.Tensor


@dataclass
class CollisionSplits:
    overlapped_lengths: torch.Tensor
    nonoverlapped_lengths: torch.Tensor


@dataclass
class CollisionPermutation:
    """``forward_permute[i]`` = position of original value i inside ``cat([overlapped rows, non-overlapped rows])``."""

    forward_permute: torch.Tensor
    num_overlapped: torch.Tensor


class CollisionHandlerBase:
    def detect_collisions(self, features: KeyedJaggedTensor, prev_keys: Optional[torch.Tensor]) -> CollisionResult:
        raise NotImplementedError


class BooleanCollisionHandler(CollisionHandlerBase):
    def __init__(self, unit_row_base: List[int], total_rows: int, device: torch.device) -> None:
        """``unit_row_base[u]``: first row key of the table shard unit u reads (units of features that share a table share a base)."""
        self._num_units = len(unit_row_base)
        self._total_rows = int(total_rows)
        self._row_base = torch.tensor(list(unit_row_base) or [0], dtype=torch.int64, device=device)
        self._seen = torch.zeros(max(self._total_rows, 1), dtype=torch.bool, device=device)
        self._device = device

    def row_keys(self, features: KeyedJaggedTensor) -> torch.Tensor:
        """unit_row_base[unit of value] + value, with the unit recovered on the device from the offsets."""
        values = features.values().long()
        n = values.numel()
        if n == 0 or self._num_units == 0:
            return values
        Bg = features.stride()
        bounds = features.offsets()[Bg::Bg][: self._num_units].long()      # end offset of every unit's segment
        unit = torch.bucketize(torch.arange(n, device=values.device), bounds, right=True).clamp_(max=self._num_units - 1)
        return self._row_base[unit] + values

    def detect_collisions(self, features: KeyedJaggedTensor, prev_keys: Optional[torch.Tensor]) -> CollisionResult:
        keys = self.row_keys(features)
        if prev_keys is None or prev_keys.numel() == 0 or keys.numel() == 0:
            return CollisionResult(torch.zeros_like(keys, dtype=torch.bool), None if prev_keys is None else torch.zeros_like(prev_keys, dtype=torch.bool), keys)
        seen = self._seen
        seen[prev_keys] = True
        fwd = seen[keys]
        seen[prev_keys] = False          # the table is all-false between calls
        seen[keys] = True
        bwd = seen[prev_keys]
        seen[keys] = False
        return CollisionResult(fwd, bwd, keys)


def create_collision_handler(checker_type: OverlappingCheckerType, unit_row_base: List[int], total_rows: int, device: torch.device) -> CollisionHandlerBase:
    if OverlappingCheckerType(checker_type) == OverlappingCheckerType.BOOLEAN:
        return BooleanCollisionHandler(unit_row_base, total_rows, device)
    raise ValueError(f"unsupported overlap checker {checker_type}")


def split_features_by_values_mask(features: KeyedJaggedTensor, mask: torch.Tensor) -> Tuple[KeyedJaggedTensor, KeyedJaggedTensor, CollisionPermutation]:
    """(overlapped KJT, non-overlapped KJT, permutation) with the bag structure of ``features`` preserved in both parts."""
    offsets = features.offsets().long()
    m = mask.to(torch.int64)
    c_ol = torch.cat([m.new_zeros(1), m.cumsum(0)])
    c_nol = torch.cat([m.new_zeros(1), (1 - m).cumsum(0)])
    len_ol = c_ol[offsets[1:]] - c_ol[offsets[:-1]]
    len_nol = c_nol[offsets[1:]] - c_nol[offsets[:-1]]
    values, weights = features.values(), features.weights_or_none()
    keep = ~mask

    def part(sel: torch.Tensor, lengths: torch.Tensor) -> KeyedJaggedTensor:
        return KeyedJaggedTensor(keys=features.keys(), values=values[sel], weights=None if weights is None else weights[sel],
                                 lengths=lengths.to(features.lengths().dtype), stride=features.stride())

    n_ol = c_ol[-1]
    perm = torch.where(mask, c_ol[1:] - 1, n_ol + c_nol[1:] - 1)
    return part(mask, len_ol), part(keep, len_nol), CollisionPermutation(perm, n_ol)
