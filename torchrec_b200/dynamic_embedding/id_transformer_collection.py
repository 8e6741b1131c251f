"""Per-table id transformers + PS clients behind one KJT -> KJT transform
(reference contrib/dynamic_embedding/.../id_transformer_collection.py, id_transformer_group.py)."""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch

from ..modules.embedding_configs import BaseEmbeddingConfig
from ..sparse.jagged_tensor import KeyedJaggedTensor
from .id_transformer import IDTransformer
from .ps import PS


class IDTransformerCollection:
    def __init__(self, tables: List[BaseEmbeddingConfig], eviction_config: Optional[dict] = None, transform_config: Optional[dict] = None,
                 ps_collection: Optional[Dict[str, PS]] = None, evict_fraction: float = 0.05) -> None:
        """tables: configs whose ``num_embeddings`` is the CACHE size (the id space itself is unbounded)."""
        self._tables = tables
        self._feature_table: Dict[str, int] = {f: i for i, t in enumerate(tables) for f in t.feature_names}
        self._transformers = [IDTransformer(t.num_embeddings, eviction_config, transform_config) for t in tables]
        self._ps = ps_collection or {}
        self._evict_fraction = evict_fraction
        self._time = 0

    def transform(self, global_features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        """Global-id KJT -> cache-id KJT. Evicts (write-back through the PS) and fetches as needed, so that when this
        returns the cache tables hold the rows of every id in the batch."""
        self._time += 1
        lpk = global_features.length_per_key()
        vals = list(torch.split(global_features.values().cpu(), lpk))
        for i, key in enumerate(global_features.keys()):
            ti = self._feature_table.get(key)
            if ti is None or vals[i].numel() == 0:
                continue
            tr, table = self._transformers[ti], self._tables[ti]
            ps = self._ps.get(table.name)
            cache_ids, ok, to_fetch = tr.transform(vals[i], self._time)
            rounds = 0
            while not ok:
                n_evict = max(int(table.num_embeddings * self._evict_fraction), int((cache_ids < 0).sum()))
                evicted = tr.evict(n_evict)
                if ps is not None:
                    # ids inserted in this very call were never materialised: nothing to write back for them
                    fresh = set(to_fetch[:, 0].tolist())
                    keep = torch.tensor([g not in fresh for g in evicted[:, 0].tolist()], dtype=torch.bool)
                    ps.evict(evicted[keep])
                    gone = set(evicted[:, 0].tolist())
                    to_fetch = to_fetch[torch.tensor([g not in gone for g in to_fetch[:, 0].tolist()], dtype=torch.bool)] if to_fetch.numel() else to_fetch
                cache_ids, ok, more = tr.transform(vals[i], self._time)
                to_fetch = torch.cat([to_fetch, more]) if more.numel() else to_fetch
                rounds += 1
                if rounds > 64:
                    raise RuntimeError(f"table {table.name}: one batch needs more distinct ids than the cache holds ({table.num_embeddings})")
            if ps is not None and to_fetch.numel():
                ps.fetch(to_fetch)
            vals[i] = cache_ids
        dev = global_features.values().device
        return KeyedJaggedTensor(keys=global_features.keys(), values=torch.cat(vals).to(dev), lengths=global_features.lengths(), weights=global_features.weights_or_none(),
                                 stride=global_features.stride(), length_per_key=lpk)

    def save(self) -> None:
        """Write every cached row back to the PS (checkpoint)."""
        for tr, table in zip(self._transformers, self._tables):
            ps = self._ps.get(table.name)
            if ps is not None:
                ps.evict(tr.save()[:, :2])
                ps.wait()

    @property
    def transformers(self) -> List[IDTransformer]:
        return self._transformers
