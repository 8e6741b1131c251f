"""Per-table store of looked-up ids (+ states) with windowed compaction
(reference torchrec/distributed/model_tracker/delta_store.py:24-341)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .types import IndexedLookup, RawIndexedLookup, UniqueRows, UpdateMode


def compute_unique_rows(ids: List[torch.Tensor], states: Optional[List[torch.Tensor]], mode: UpdateMode) -> UniqueRows:
    """Unique ids of a list of lookups; the state kept per id is the FIRST or LAST one seen."""
    cat_ids = torch.cat(ids)
    if states is None or mode == UpdateMode.NONE:
        return UniqueRows(ids=torch.unique(cat_ids), states=None)
    cat_states = torch.cat(states)
    uniq, inv = torch.unique(cat_ids, return_inverse=True)
    pos = torch.arange(cat_ids.numel(), device=cat_ids.device)
    pick = torch.full((uniq.numel(),), cat_ids.numel() if mode == UpdateMode.FIRST else -1, dtype=torch.long, device=cat_ids.device)
    pick.scatter_reduce_(0, inv, pos, reduce="amin" if mode == UpdateMode.FIRST else "amax", include_self=True)
    return UniqueRows(ids=uniq, states=cat_states[pick])


class DeltaStore:
    def __init__(self, updateMode: UpdateMode = UpdateMode.NONE) -> None:
        self.updateMode = updateMode
        self.per_fqn_lookups: Dict[str, List[IndexedLookup]] = {}

    def append(self, batch_idx: int, fqn: str, ids: torch.Tensor, states: Optional[torch.Tensor]) -> None:
        self.per_fqn_lookups.setdefault(fqn, []).append(IndexedLookup(batch_idx=batch_idx, ids=ids, states=states))

    def delete(self, up_to_idx: Optional[int] = None) -> None:
        if up_to_idx is None:
            self.per_fqn_lookups = {}
            return
        for fqn, lookups in list(self.per_fqn_lookups.items()):
            self.per_fqn_lookups[fqn] = [lk for lk in lookups if lk.batch_idx >= up_to_idx]

    def compact(self, start_idx: int, end_idx: int) -> None:
        """Merge the lookups of batches [start_idx, end_idx) into one unique entry at ``start_idx``."""
        assert start_idx < end_idx, f"start_idx {start_idx} must be smaller than end_idx {end_idx}"
        for fqn, lookups in list(self.per_fqn_lookups.items()):
            window = [lk for lk in lookups if start_idx <= lk.batch_idx < end_idx]
            if len(window) <= 1:
                continue
            rows = compute_unique_rows([lk.ids for lk in window], [lk.states for lk in window] if window[0].states is not None else None, self.updateMode)
            merged = IndexedLookup(batch_idx=start_idx, ids=rows.ids, states=rows.states, compact=True)
            out, placed = [], False
            for lk in lookups:
                if start_idx <= lk.batch_idx < end_idx:
                    if not placed:
                        out.append(merged)
                        placed = True
                else:
                    out.append(lk)
            self.per_fqn_lookups[fqn] = out

    def get_indexed_lookups(self, start_idx: int, end_idx: int) -> Dict[str, List[IndexedLookup]]:
        return {fqn: [lk for lk in lookups if start_idx <= lk.batch_idx < end_idx] for fqn, lookups in self.per_fqn_lookups.items()}

    def get_unique(self, from_idx: int = 0) -> Dict[str, UniqueRows]:
        res: Dict[str, UniqueRows] = {}
        for fqn, lookups in self.per_fqn_lookups.items():
            sel = [lk for lk in lookups if lk.batch_idx >= from_idx]
            if sel:
                res[fqn] = compute_unique_rows([lk.ids for lk in sel], [lk.states for lk in sel] if sel[0].states is not None else None, self.updateMode)
        return res


DeltaStoreTrec = DeltaStore


class RawIdTrackerStore:
    """Raw (pre-remap) ids of managed-collision tables, for MPZCH id streaming."""

    def __init__(self, updateMode: UpdateMode = UpdateMode.NONE) -> None:
        self.per_fqn_lookups: Dict[str, List[RawIndexedLookup]] = {}

    def append(self, batch_idx: int, fqn: str, ids: torch.Tensor, raw_ids: Optional[torch.Tensor] = None, runtime_meta: Optional[torch.Tensor] = None) -> None:
        self.per_fqn_lookups.setdefault(fqn, []).append(RawIndexedLookup(batch_idx, ids, raw_ids, runtime_meta))

    def delete(self, up_to_idx: Optional[int] = None) -> None:
        if up_to_idx is None:
            self.per_fqn_lookups = {}
        else:
            self.per_fqn_lookups = {k: [lk for lk in v if lk.batch_idx >= up_to_idx] for k, v in self.per_fqn_lookups.items()}

    def compact(self, start_idx: int, end_idx: int) -> None:
        pass

    def get_indexed_lookups(self, start_idx: int, end_idx: int) -> Dict[str, List[RawIndexedLookup]]:
        return {k: [lk for lk in v if start_idx <= lk.batch_idx < end_idx] for k, v in self.per_fqn_lookups.items()}

    def get_unique(self, from_idx: int = 0) -> Dict[str, UniqueRows]:
        return {k: UniqueRows(torch.unique(torch.cat([lk.ids for lk in v if lk.batch_idx >= from_idx])), None) for k, v in self.per_fqn_lookups.items() if v}
