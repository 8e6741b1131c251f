"""Raw-id tracker: remembers, per managed-collision table, which RAW ids were mapped to which table rows since the last read.

Reference: ``torchrec/distributed/model_tracker/trackers/raw_id_tracker.py:40`` (``RawIdTracker``) - used to stream (raw id -> row) pairs of MPZCH tables to
an inference side that keeps its own id map. The reference threads callbacks through the embedding kernels (``init_raw_id_tracker``); here every
managed-collision module is observed with a forward hook: its input carries the raw ids, its output the remapped rows, aligned one to one.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
from torch import nn

from ..delta_store import RawIdTrackerStore
from ..model_delta_tracker import ModelDeltaTracker
from ..types import RawIndexedLookup, UniqueRows


class RawIdTracker(ModelDeltaTracker):
    def __init__(self, model: nn.Module, consumers: Optional[List[str]] = None, delete_on_read: bool = True, fqns_to_skip: Iterable[str] = (),
                 tables_to_track: Optional[Iterable[str]] = None) -> None:
        self._model = model
        self._consumers = consumers or [self.DEFAULT_CONSUMER]
        self._delete_on_read = delete_on_read
        self._fqns_to_skip = list(fqns_to_skip)
        self._tables = set(tables_to_track) if tables_to_track is not None else None
        self.curr_batch_idx = 0
        self.per_consumer_batch_idx: Dict[str, int] = {c: 0 for c in self._consumers}
        self.store = RawIdTrackerStore()
        self.tracked_modules: Dict[str, nn.Module] = {}
        self.table_to_fqn: Dict[str, str] = {}
        self._handles: List[torch.utils.hooks.RemovableHandle] = []
        self._install()

    # ---- discovery: every ManagedCollisionModule reachable from the model, keyed by its table name --------------------------
    def _should_skip_fqn(self, fqn: str) -> bool:
        return any(s in fqn for s in self._fqns_to_skip)

    def _should_track_table(self, table: str) -> bool:
        return self._tables is None or table in self._tables

    def fqn_to_feature_names(self) -> Dict[str, List[str]]:
        return {fqn: list(getattr(m, "_feature_names", []) or []) for fqn, m in self.tracked_modules.items()}

    def _install(self) -> None:
        from ....modules.mc_modules import ManagedCollisionModule

        for fqn, m in self._model.named_modules():
            holder = getattr(m, "_managed_collision_modules", None)
            if holder is None or self._should_skip_fqn(fqn):
                continue
            for table, mc in holder.items():
                if not isinstance(mc, ManagedCollisionModule) or not self._should_track_table(table):
                    continue
                clean = fqn.replace("_dmp_wrapped_module.", "").replace("module.", "")
                tfqn = f"{clean}._managed_collision_modules.{table}" if clean else f"_managed_collision_modules.{table}"
                if tfqn in self.tracked_modules:
                    continue
                self.tracked_modules[tfqn] = mc
                self.table_to_fqn[table] = tfqn
                self._handles.append(mc.register_forward_hook(self._make_hook(tfqn)))

    def _make_hook(self, tfqn: str):
        def hook(module: nn.Module, args, output) -> None:
            if not module.training and not getattr(self, "track_in_eval", False):
                return
            raw = args[0]
            with torch.no_grad():
                for feat, jt in output.items():
                    if feat in raw:
                        self.record_raw(tfqn, jt.values().detach(), raw[feat].values().detach())

        return hook

    # ---- recording -----------------------------------------------------------------------------------------------------------
    def record_raw(self, fqn: str, ids: torch.Tensor, raw_ids: torch.Tensor, runtime_meta: Optional[torch.Tensor] = None) -> None:
        assert ids.numel() == raw_ids.numel(), f"{fqn}: {ids.numel()} remapped ids for {raw_ids.numel()} raw ids"
        self.store.append(self.curr_batch_idx, fqn, ids.clone(), raw_ids.clone(), runtime_meta)

    def record_lookup(self, emb_module: nn.Module, kjt, states: Optional[torch.Tensor] = None, raw_ids: Optional[torch.Tensor] = None) -> None:
        fqn = next((f for f, m in self.tracked_modules.items() if m is emb_module), type(emb_module).__name__)
        vals = kjt.values() if hasattr(kjt, "values") and callable(kjt.values) else kjt
        self.record_raw(fqn, vals.detach(), (raw_ids if raw_ids is not None else vals).detach())

    def step(self) -> None:
        self.curr_batch_idx += 1

    # ---- reading ---------------------------------------------------------------------------------------------------------------
    def _window(self, consumer: Optional[str]):
        consumer = consumer or self.DEFAULT_CONSUMER
        assert consumer in self.per_consumer_batch_idx, f"unknown consumer {consumer}"
        start, end = self.per_consumer_batch_idx[consumer], self.curr_batch_idx + 1
        return consumer, start, end

    def get_indexed_lookups(self, consumer: Optional[str] = None) -> Dict[str, List[RawIndexedLookup]]:
        """All (rows, raw ids) batches since this consumer's last read, in arrival order."""
        consumer, start, end = self._window(consumer)
        out = self.store.get_indexed_lookups(start, end)
        self._advance(consumer, end)
        return out

    def get_raw_id_map(self, consumer: Optional[str] = None) -> Dict[str, Dict[str, torch.Tensor]]:
        """Per table the LATEST raw id of every touched row: ``{"ids": rows (sorted, unique), "raw_ids": raw id now living in that row}``."""
        out: Dict[str, Dict[str, torch.Tensor]] = {}
        for fqn, lookups in self.get_indexed_lookups(consumer).items():
            if not lookups:
                continue
            ids = torch.cat([lk.ids.reshape(-1) for lk in lookups])
            raw = torch.cat([lk.raw_ids.reshape(-1) for lk in lookups])
            # last write wins: scan from the back, keep the first occurrence of every row
            rev_ids, rev_raw = ids.flip(0), raw.flip(0)
            uniq, inv = torch.unique(rev_ids, return_inverse=True)
            first = torch.full((uniq.numel(),), rev_ids.numel(), dtype=torch.long, device=ids.device)
            first.scatter_reduce_(0, inv, torch.arange(rev_ids.numel(), device=ids.device), reduce="amin")
            out[fqn] = {"ids": uniq, "raw_ids": rev_raw[first]}
        return out

    def get_unique_ids(self, consumer: Optional[str] = None) -> Dict[str, torch.Tensor]:
        return {fqn: m["ids"] for fqn, m in self.get_raw_id_map(consumer).items()}

    def get_unique(self, consumer: Optional[str] = None, top_percentage: Optional[float] = 1.0, per_table_percentage=None, sorted_by_indices: Optional[bool] = True) -> Dict[str, UniqueRows]:
        return {fqn: UniqueRows(m["ids"], m["raw_ids"]) for fqn, m in self.get_raw_id_map(consumer).items()}

    def _advance(self, consumer: str, end: int) -> None:
        self.per_consumer_batch_idx[consumer] = end
        if self._delete_on_read:
            self.store.delete(up_to_idx=min(self.per_consumer_batch_idx.values()))

    def clear(self, consumer: Optional[str] = None) -> None:
        if consumer is None:
            self.store.delete()
            self.per_consumer_batch_idx = {c: self.curr_batch_idx for c in self._consumers}
        else:
            self._advance(consumer, self.curr_batch_idx + 1)

    def get_tracked_modules(self) -> Dict[str, nn.Module]:
        return self.tracked_modules

    def remove_hooks(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
