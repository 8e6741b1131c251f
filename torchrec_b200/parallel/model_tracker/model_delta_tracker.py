"""Model delta tracker: which embedding rows did training touch since a consumer last asked?
(reference torchrec/distributed/model_tracker/model_delta_tracker.py:66-653).

Hooks into every sharded embedding collection's lookup engine: after the input dist, the ids this rank looks up are
recorded per table (as global row ids), optionally with the row's embedding / optimizer state. Consumers (a publisher of
fresh rows to inference, a delta checkpointer) call ``get_unique`` and receive what changed since their previous call."""
from __future__ import annotations

import logging
from abc import ABC, abstractmethod
from typing import Dict, Iterable, List, Optional

import torch
from torch import nn

from .delta_store import DeltaStore
from .types import TrackingMode, UniqueRows, UpdateMode

logger = logging.getLogger(__name__)

UPDATE_MODE_MAP: Dict[TrackingMode, UpdateMode] = {
    TrackingMode.ID_ONLY: UpdateMode.NONE,
    TrackingMode.EMBEDDING: UpdateMode.FIRST,
    TrackingMode.MOMENTUM_LAST: UpdateMode.LAST,
    TrackingMode.MOMENTUM_DIFF: UpdateMode.FIRST,
    TrackingMode.ROWWISE_ADAGRAD: UpdateMode.LAST,
}


class ModelDeltaTracker(ABC):
    DEFAULT_CONSUMER: str = "default"

    @abstractmethod
    def record_lookup(self, emb_module: nn.Module, kjt, states: torch.Tensor) -> None:
        ...

    @abstractmethod
    def get_unique_ids(self, consumer: Optional[str] = None) -> Dict[str, torch.Tensor]:
        ...

    @abstractmethod
    def get_unique(self, consumer: Optional[str] = None, top_percentage: Optional[float] = 1.0, per_table_percentage=None, sorted_by_indices: Optional[bool] = True) -> Dict[str, UniqueRows]:
        ...

    @abstractmethod
    def clear(self, consumer: Optional[str] = None) -> None:
        ...

    @abstractmethod
    def step(self) -> None:
        ...


class ModelDeltaTrackerTrec(ModelDeltaTracker):
    def __init__(self, model: nn.Module, consumers: Optional[List[str]] = None, delete_on_read: bool = True, auto_compact: bool = False,
                 mode: TrackingMode = TrackingMode.ID_ONLY, fqns_to_skip: Iterable[str] = ()) -> None:
        self._model = model
        self._consumers = consumers or [self.DEFAULT_CONSUMER]
        self._delete_on_read = delete_on_read
        self._auto_compact = auto_compact
        self._mode = mode
        self._fqns_to_skip = list(fqns_to_skip)
        self.per_consumer_batch_idx: Dict[str, int] = {c: -1 for c in self._consumers}
        self.curr_batch_idx = 0
        self.curr_compact_index = 0
        self.store = DeltaStore(UPDATE_MODE_MAP[mode])
        self.tracked_modules: Dict[str, nn.Module] = {}
        self.table_to_fqn: Dict[str, str] = {}
        self.feature_to_fqn: Dict[str, str] = {}
        self._fqn_to_feature_map: Dict[str, List[str]] = {}
        self._handles = []
        self.fqn_to_feature_names()
        self._install_hooks()

    # ---- discovery --------------------------------------------------------------------------------------------
    def _clean_fqn_fn(self, fqn: str) -> str:
        for junk in ("_dmp_wrapped_module.", "module."):
            fqn = fqn.replace(junk, "")
        return fqn

    def fqn_to_feature_names(self) -> Dict[str, List[str]]:
        if self._fqn_to_feature_map:
            return self._fqn_to_feature_map
        from ..embedding import ShardedEmbeddingCollection
        from ..embeddingbag import ShardedEmbeddingBagCollection

        for fqn, m in self._model.named_modules():
            if isinstance(m, (ShardedEmbeddingBagCollection, ShardedEmbeddingCollection)):
                fqn = self._clean_fqn_fn(fqn)
                if any(skip in fqn for skip in self._fqns_to_skip):
                    continue
                self.tracked_modules[fqn] = m
                kind = "embedding_bags" if isinstance(m, ShardedEmbeddingBagCollection) else "embeddings"
                cfgs = m.embedding_bag_configs() if isinstance(m, ShardedEmbeddingBagCollection) else m.embedding_configs()
                for c in cfgs:
                    tfqn = f"{fqn}.{kind}.{c.name}" if fqn else f"{kind}.{c.name}"
                    self.table_to_fqn[c.name] = tfqn
                    self._fqn_to_feature_map[tfqn] = list(c.feature_names)
                    for f in c.feature_names:
                        self.feature_to_fqn.setdefault(f, tfqn)
        return self._fqn_to_feature_map

    def get_tracked_modules(self) -> Dict[str, nn.Module]:
        return self.tracked_modules

    def _install_hooks(self) -> None:
        for fqn, m in self.tracked_modules.items():
            eng = m.engine if hasattr(m, "engine") else getattr(m, "_engine", None)
            if eng is None:
                continue
            if self._mode in (TrackingMode.MOMENTUM_LAST, TrackingMode.MOMENTUM_DIFF, TrackingMode.ROWWISE_ADAGRAD):
                for shard, _w, st, _t in eng.local_shard_views():
                    assert "momentum1" in st or len(st) > 0, f"{self._mode} needs a fused optimizer with state on table {shard.name}"
            self._handles.append(eng.register_lookup_hook(self._on_lookup))

    # ---- recording -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _on_lookup(self, eng, dist_features) -> None:
        Bg = dist_features.stride()
        offsets = dist_features.offsets()
        values = dist_features.values()
        views = {(s.name, s.row_off, s.col_off): (w, st) for s, w, st, _ in eng.local_shard_views()}
        seen = set()
        bounds = offsets[torch.arange(0, len(eng.local_units) + 1, device=offsets.device) * Bg].tolist() if len(eng.local_units) else []
        for k, u in enumerate(eng.local_units):
            s = u.shard
            if (s.name, s.row_off, u.feature) in seen:  # further column shards of the same rows
                continue
            seen.add((s.name, s.row_off, u.feature))
            lo, hi = int(bounds[k]), int(bounds[k + 1])
            if hi <= lo:
                continue
            local = values[lo:hi].long()
            fqn = self.table_to_fqn.get(s.name)
            if fqn is None:
                continue
            states = None
            w, st = views[(s.name, s.row_off, s.col_off)]
            if self._mode == TrackingMode.EMBEDDING:
                states = w[local].detach().clone()
            elif self._mode in (TrackingMode.MOMENTUM_LAST, TrackingMode.MOMENTUM_DIFF, TrackingMode.ROWWISE_ADAGRAD):
                m1 = st.get("momentum1", next(iter(st.values())))
                states = m1[local].detach().clone()
                if states.dim() == 1:
                    states = states.unsqueeze(1)
            self.store.append(self.curr_batch_idx, fqn, local + s.row_off, states)

    def record_lookup(self, emb_module: nn.Module, kjt, states: Optional[torch.Tensor] = None) -> None:
        """Manual recording of a (global-id) KJT, for modules without an engine hook."""
        lpk = kjt.length_per_key()
        vals = torch.split(kjt.values(), lpk)
        st = torch.split(states, lpk) if states is not None else [None] * len(lpk)
        for k, v, s_ in zip(kjt.keys(), vals, st):
            fqn = self.feature_to_fqn.get(k)
            if fqn is not None and v.numel():
                self.store.append(self.curr_batch_idx, fqn, v.long(), s_)

    def record_ids(self, kjt) -> None:
        self.record_lookup(None, kjt, None)

    def record_embeddings(self, emb_module: nn.Module, kjt, states: torch.Tensor) -> None:
        """EMBEDDING mode by hand: ``states`` = the embedding rows looked up for ``kjt`` (one row per id, in value order)."""
        assert states is not None and states.shape[0] == kjt.values().numel(), "one state row per id"
        self.record_lookup(emb_module, kjt, states.detach())

    def _state_rows_of(self, emb_module: nn.Module, kjt, state_name: str) -> torch.Tensor:
        """Rows of an optimizer state for the (global) ids of ``kjt``, read from the tables' full state tensors of the module's fused
        optimizer (``<table>.<state_name>`` or ``<table>.weight.<state_name>`` keys)."""
        opt = getattr(emb_module, "fused_optimizer", None)
        assert opt is not None, "recording optimizer state needs a module with a fused optimizer"
        state = opt.state_dict()["state"]
        cfgs = emb_module.embedding_bag_configs() if hasattr(emb_module, "embedding_bag_configs") else emb_module.embedding_configs()
        table_of = {f: c.name for c in cfgs for f in c.feature_names}
        parts = []
        for k, v in zip(kjt.keys(), torch.split(kjt.values(), kjt.length_per_key())):
            t = table_of[k]
            entry = next((val for key, val in state.items() if key.endswith(f"{t}.weight") or key.endswith(t)), None)
            assert entry is not None and state_name in entry, f"no optimizer state {state_name!r} for table {t}"
            full = entry[state_name]
            full = full.full_tensor() if hasattr(full, "full_tensor") else (full.local_tensor() if hasattr(full, "local_tensor") and not isinstance(full, torch.Tensor) else full)
            rows = full[v.long().to(full.device)]
            parts.append(rows.unsqueeze(1) if rows.dim() == 1 else rows)
        return torch.cat(parts) if parts else torch.zeros(0, 1)

    def record_momentum(self, emb_module: nn.Module, kjt) -> None:
        """MOMENTUM_LAST / MOMENTUM_DIFF by hand: first-moment rows of the ids in ``kjt`` (Adam / LAMB families)."""
        self.record_lookup(emb_module, kjt, self._state_rows_of(emb_module, kjt, "momentum1"))

    def record_rowwise_optim_state(self, emb_module: nn.Module, kjt) -> None:
        """ROWWISE_ADAGRAD by hand: the per-row accumulator (one float per row) of the ids in ``kjt``."""
        self.record_lookup(emb_module, kjt, self._state_rows_of(emb_module, kjt, "momentum1"))

    # ---- reading ---------------------------------------------------------------------------------------------
    def step(self) -> None:
        self.curr_batch_idx += 1
        if self._auto_compact:
            self.trigger_compaction()

    def trigger_compaction(self) -> None:
        if self.curr_compact_index >= self.curr_batch_idx:
            return
        start = max(self.per_consumer_batch_idx.values())
        start = max(start, self.curr_compact_index, 0)
        if start < self.curr_batch_idx:
            self.compact(start, self.curr_batch_idx)
        self.curr_compact_index = self.curr_batch_idx

    def compact(self, start_idx: int, end_idx: int) -> None:
        self.store.compact(start_idx, end_idx)

    def get_unique_ids(self, consumer: Optional[str] = None) -> Dict[str, torch.Tensor]:
        return {k: v.ids for k, v in self.get_unique(consumer).items()}

    def get_unique(self, consumer: Optional[str] = None, top_percentage: Optional[float] = 1.0, per_table_percentage: Optional[Dict[str, float]] = None,
                   sorted_by_indices: Optional[bool] = True) -> Dict[str, UniqueRows]:
        consumer = consumer or self.DEFAULT_CONSUMER
        assert consumer in self.per_consumer_batch_idx, f"consumer {consumer} not registered"
        from_idx = max(self.per_consumer_batch_idx[consumer], 0)
        rows = self.store.get_unique(from_idx)
        if self._mode == TrackingMode.MOMENTUM_DIFF:
            latest = self._latest_states(rows)
            rows = {k: UniqueRows(v.ids, (latest[k] - v.states) if v.states is not None else None) for k, v in rows.items()}
        out: Dict[str, UniqueRows] = {}
        for fqn, r in rows.items():
            pct = (per_table_percentage or {}).get(fqn.rsplit(".", 1)[-1], top_percentage if top_percentage is not None else 1.0)
            if pct < 1.0 and r.states is not None and r.ids.numel():
                k = max(1, int(r.ids.numel() * pct))
                score = r.states.float().abs().sum(1) if r.states.dim() > 1 else r.states.float().abs()
                top = torch.topk(score, k).indices
                if sorted_by_indices:
                    top = top[torch.argsort(r.ids[top])]
                r = UniqueRows(r.ids[top], r.states[top])
            out[fqn] = r
        self.per_consumer_batch_idx[consumer] = self.curr_batch_idx
        if self._delete_on_read:
            self.store.delete(up_to_idx=min(self.per_consumer_batch_idx.values()))
        return out

    get_delta = get_unique

    def get_delta_ids(self, consumer: Optional[str] = None) -> Dict[str, torch.Tensor]:
        return self.get_unique_ids(consumer)

    def _latest_states(self, rows: Dict[str, UniqueRows]) -> Dict[str, torch.Tensor]:
        res: Dict[str, torch.Tensor] = {}
        for fqn, r in rows.items():
            table = fqn.rsplit(".", 1)[-1]
            mod = next(m for f, m in self.tracked_modules.items() if fqn.startswith(f))
            acc = None
            for s, w, st, _ in mod.engine.local_shard_views():
                if s.name != table:
                    continue
                m1 = st.get("momentum1", next(iter(st.values())))
                sel = (r.ids >= s.row_off) & (r.ids < s.row_off + s.rows)
                if acc is None:
                    acc = torch.zeros(r.ids.numel(), *(m1.shape[1:] if m1.dim() > 1 else (1,)), device=m1.device, dtype=m1.dtype)
                v = m1[r.ids[sel] - s.row_off]
                acc[sel] = v if v.dim() > 1 else v.unsqueeze(1)
            res[fqn] = acc
        return res

    def get_latest(self) -> Dict[str, torch.Tensor]:
        """Current embedding rows of every id touched since tracking began (per table FQN)."""
        res: Dict[str, torch.Tensor] = {}
        for fqn, r in self.store.get_unique(0).items():
            table = fqn.rsplit(".", 1)[-1]
            mod = next(m for f, m in self.tracked_modules.items() if fqn.startswith(f))
            for s, w, _st, _ in mod.engine.local_shard_views():
                if s.name == table:
                    sel = (r.ids >= s.row_off) & (r.ids < s.row_off + s.rows)
                    res[fqn] = w[r.ids[sel] - s.row_off].detach().clone()
                    break
        return res

    def clear(self, consumer: Optional[str] = None) -> None:
        if consumer is None:
            self.store.delete()
            for c in self.per_consumer_batch_idx:
                self.per_consumer_batch_idx[c] = self.curr_batch_idx
        else:
            self.per_consumer_batch_idx[consumer] = self.curr_batch_idx
            self.store.delete(up_to_idx=min(self.per_consumer_batch_idx.values()))

    def remove_hooks(self) -> None:
        for h in self._handles:
            h()
        self._handles = []
