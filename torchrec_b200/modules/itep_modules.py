"""In-training embedding pruning (ITEP) (reference torchrec/modules/itep_modules.py:78-700, itep_embedding_modules.py).

Tracks how often every logical row is accessed; every ``pruning_interval`` iterations the least-used rows lose their
physical row: an *address lookup* table maps logical row -> physical row, pruned rows are re-pointed to a shared slot and
their physical rows are handed to newly hot logical rows. The embedding table itself can then be smaller than the id space
(``num_embeddings_post_pruning``)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor
from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection


class GenericITEPModule(nn.Module):
    def __init__(self, table_name_to_unpruned_hash_sizes: Dict[str, int], lookups: Optional[List[nn.Module]] = None, enable_pruning: bool = True,
                 pruning_interval: int = 1001, table_name_to_pruned_hash_sizes: Optional[Dict[str, int]] = None, feature_to_table: Optional[Dict[str, str]] = None,
                 device: Optional[torch.device] = None) -> None:
        super().__init__()
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
            keep = torch.topk(util, pruned - 1).indices if pruned > 1 else util.new_zeros(0, dtype=torch.long)
            keep_mask = torch.zeros_like(util, dtype=torch.bool)
            keep_mask[keep] = True
            owns = torch.zeros_like(util, dtype=torch.bool)
            valid_owner = owner[: pruned - 1] >= 0
            owns[owner[: pruned - 1][valid_owner]] = True
            stay = keep_mask & owns  # hot rows that already own a physical row keep it
            newcomers = (keep_mask & ~owns).nonzero(as_tuple=True)[0]
            freed_phys = (~stay[owner[: pruned - 1].clamp(min=0)] | ~valid_owner).nonzero(as_tuple=True)[0]
            n = min(newcomers.numel(), freed_phys.numel())
            if n:
                addr[owner[freed_phys[:n]].clamp(min=0)] = pruned - 1  # old owners fall back to the shared slot
                owner[freed_phys[:n]] = newcomers[:n]
                addr[newcomers[:n]] = freed_phys[:n]
                changed[t] = freed_phys[:n]
            util.mul_(0.5)  # exponential decay of the access statistics
        return changed


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
