"""In-training embedding pruning (ITEP) (reference torchrec/modules/itep_modules.py:78-700, itep_embedding_modules.py).

Tracks how often every logical row is accessed; every ``pruning_interval`` iterations the least-used rows lose their
physical row: an *address lookup* table maps logical row -> physical row, pruned rows are re-pointed to a shared slot and
their physical rows are handed to newly hot logical rows. The embedding table itself can then be smaller than the id space
(``num_embeddings_post_pruning``)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor
from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection


class GenericITEPModule(nn.Module):
    def __init__(self, table_name_to_unpruned_hash_sizes: Dict[str, int], lookups: Optional[List[nn.Module]] = None, enable_pruning: bool = True,
                 pruning_interval: int = 1001, table_name_to_pruned_hash_sizes: Optional[Dict[str, int]] = None, feature_to_table: Optional[Dict[str, str]] = None,
                 device: Optional[torch.device] = None, pg: Optional[Any] = None, table_name_to_sharding_type: Optional[Dict[str, str]] = None,
                 pruning_logger_type: Optional[Any] = None) -> None:
        super().__init__()
        self.pg = pg  # process group the pruning decisions are agreed over (row-wise sharded tables)
        self.table_name_to_sharding_type: Dict[str, str] = dict(table_name_to_sharding_type or {})
        if pruning_logger_type is None:
            from .pruning_logger import PruningLoggerDefault as pruning_logger_type  # noqa: N813
        self.pruning_logger = pruning_logger_type
        self.enable_pruning = enable_pruning
        self.pruning_interval = pruning_interval
        self.table_name_to_unpruned_hash_sizes = table_name_to_unpruned_hash_sizes
        self.table_name_to_pruned_hash_sizes = table_name_to_pruned_hash_sizes or {}
        self.feature_to_table = feature_to_table or {}
        self.register_buffer("iter", torch.tensor(0, dtype=torch.int64))
        self._tables = [t for t in table_name_to_unpruned_hash_sizes if t in self.table_name_to_pruned_hash_sizes]
        for t in self._tables:
            unpruned, pruned = table_name_to_unpruned_hash_sizes[t], self.table_name_to_pruned_hash_sizes[t]
            assert pruned <= unpruned
            # logical row -> physical row; initially the first `pruned` rows are mapped 1:1, the rest share the last slot
            addr = torch.arange(unpruned, dtype=torch.int64, device=device).clamp(max=pruned - 1)
            self.register_buffer(f"address_lookup_{t}", addr)
            self.register_buffer(f"row_util_{t}", torch.zeros(unpruned, dtype=torch.float32, device=device))
            self.register_buffer(f"owner_{t}", torch.arange(pruned, dtype=torch.int64, device=device))  # physical row -> logical owner
        self.last_pruned_iter = -1
        self.last_changed: Dict[str, torch.Tensor] = {}     # physical rows handed to a new logical owner by the latest prune()
        self.evicted_total: Dict[str, int] = {t: 0 for t in self._tables}
        self.num_prunes = 0

    def _addr(self, t: str) -> torch.Tensor:
        return getattr(self, f"address_lookup_{t}")

    def _util(self, t: str) -> torch.Tensor:
        return getattr(self, f"row_util_{t}")

    @torch.no_grad()
    def forward(self, sparse_features: KeyedJaggedTensor, cur_iter: int) -> KeyedJaggedTensor:
        if not self.enable_pruning or not self._tables:
            return sparse_features
        values = sparse_features.values()
        lpk = sparse_features.length_per_key()
        parts = list(torch.split(values, lpk))
        for i, key in enumerate(sparse_features.keys()):
            t = self.feature_to_table.get(key)
            if t not in self._tables:
                continue
            v = parts[i].long()
            if self.training:
                self._util(t).index_add_(0, v, torch.ones_like(v, dtype=torch.float32))
            parts[i] = self._addr(t)[v].to(values.dtype)
        if self.training and cur_iter > 0 and cur_iter % self.pruning_interval == 0 and cur_iter != self.last_pruned_iter:
            self.prune()
            self.last_pruned_iter = cur_iter
        return KeyedJaggedTensor(keys=sparse_features.keys(), values=torch.cat(parts), lengths=sparse_features.lengths(), weights=sparse_features.weights_or_none(),
                                 stride=sparse_features.stride(), length_per_key=lpk)

    @torch.no_grad()
    def prune(self) -> Dict[str, torch.Tensor]:
        """Give physical rows to the hottest logical rows. Returns, per table, the physical rows whose owner changed (to be re-initialised)."""
        changed: Dict[str, torch.Tensor] = {}
        for t in self._tables:
            pruned = self.table_name_to_pruned_hash_sizes[t]
            util, addr, owner = self._util(t), self._addr(t), getattr(self, f"owner_{t}")
            own = owner[: pruned - 1]
            valid_owner = own >= 0
            owns = torch.zeros_like(util, dtype=torch.bool)
            owns[own[valid_owner]] = True
            # candidates: accessed logical rows without a physical row, hottest first; victims: physical rows whose owner is coldest
            # first (an unowned physical row counts as utilisation -1). A candidate only displaces a strictly colder owner, so rows
            # that were never accessed do not shuffle owners around.
            cand = ((util > 0) & ~owns).nonzero(as_tuple=True)[0]
            changed_rows = util.new_zeros(0, dtype=torch.long)
            if cand.numel() and pruned > 1:
                cand = cand[torch.argsort(util[cand], descending=True)][: pruned - 1]
                owner_util = torch.where(valid_owner, util[own.clamp(min=0)], torch.full_like(util[: pruned - 1], -1.0))
                victims = torch.argsort(owner_util)[: cand.numel()]
                take = util[cand[: victims.numel()]] > owner_util[victims]
                n = int(take.sum())  # (both lists are sorted: once a candidate is not hotter than its victim, none of the rest is)
                if n:
                    victims, newcomers = victims[:n], cand[:n]
                    old = own[victims]
                    addr[old[old >= 0]] = pruned - 1  # old owners fall back to the shared slot
                    owner[victims] = newcomers
                    addr[newcomers] = victims
                    changed_rows = victims
            if changed_rows.numel():
                changed[t] = changed_rows
            util.mul_(0.5)  # exponential decay of the access statistics
        self.last_changed = changed
        self.num_prunes += 1
        for t, rows in changed.items():
            self.evicted_total[t] += int(rows.numel())
        return changed

    def eviction_stats(self) -> Dict[str, Dict[str, float]]:
        """Per table: physical rows re-assigned by the last pruning step / in total, the share of the id space that currently owns a
        physical row, and how concentrated the (decayed) access counts are (share of accesses that hit rows with a physical row) - the
        quantity the pruned table size should be tuned against (reference ``print_itep_eviction_stats`` :170-257)."""
        out: Dict[str, Dict[str, float]] = {}
        for t in self._tables:
            pruned, unpruned = self.table_name_to_pruned_hash_sizes[t], self.table_name_to_unpruned_hash_sizes[t]
            util, owner = self._util(t), getattr(self, f"owner_{t}")
            owned = owner[: pruned - 1]
            owned = owned[owned >= 0]
            total = float(util.sum())
            out[t] = {"last_evicted": float(self.last_changed[t].numel()) if t in self.last_changed else 0.0, "evicted_total": float(self.evicted_total[t]),
                      "physical_rows": float(pruned), "logical_rows": float(unpruned), "resident_fraction": float(owned.numel()) / max(unpruned, 1),
                      "access_share_resident": (float(util[owned].sum()) / total) if total > 0 else 0.0, "prunes": float(self.num_prunes)}
        return out

    @torch.no_grad()
    def reset_weight_momentum(self, collection: nn.Module, changed: Optional[Dict[str, torch.Tensor]] = None) -> int:
        """Re-initialise the physical rows that just changed owner (the new logical row must not inherit the old one's embedding) and
        clear their optimizer state when the collection keeps it next to the weights (fused collections: ``reset_rows``). Returns the
        number of rows reset (reference ``reset_weight_momentum`` :412-452)."""
        changed = self.last_changed if changed is None else changed
        n = 0
        for t, rows in changed.items():
            if rows.numel() == 0:
                continue
            if hasattr(collection, "reset_rows"):
                collection.reset_rows(t, rows)
            else:
                holder = getattr(collection, "embedding_bags", None) or getattr(collection, "embeddings", None)
                if holder is None or t not in holder:
                    continue
                w = holder[t].weight
                b = (1.0 / max(self.table_name_to_unpruned_hash_sizes[t], 1)) ** 0.5
                w.data[rows.to(w.device)] = torch.empty(rows.numel(), w.shape[1], dtype=w.dtype, device=w.device).uniform_(-b, b)
            n += int(rows.numel())
        return n


class RowwiseShardedITEPModule(GenericITEPModule):
    """ITEP for row-wise sharded tables: every rank prunes inside its own row block (no cross-rank traffic)."""

    def __init__(self, table_name_to_sharding_type: Optional[Dict[str, str]] = None, **kwargs) -> None:
        super().__init__(**kwargs)
        self.table_name_to_sharding_type = table_name_to_sharding_type or {}


# ---- moved to ``itep_embedding_modules.py`` (their reference import path); still importable from here ----
_MOVED_TO_ITEP_EMBEDDING_MODULES = ('ITEPEmbeddingBagCollection', 'ITEPEmbeddingCollection')


def __getattr__(name: str):
    if name in _MOVED_TO_ITEP_EMBEDDING_MODULES:
        from . import itep_embedding_modules as _m

        return getattr(_m, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
