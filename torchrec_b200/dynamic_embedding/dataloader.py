"""Wrap a dataloader so that id transformation (+ PS traffic) runs ahead of training on a worker thread
(reference contrib/dynamic_embedding/.../dataloader.py ``wrap``)."""
from __future__ import annotations

import queue
import threading
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional

import torch
from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor
from .id_transformer_collection import IDTransformerCollection
from .ps import PS


def _find_collections(module: nn.Module):
    from ..modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
    from ..modules.fused_embedding_modules import FusedEmbeddingBagCollection, FusedEmbeddingCollection
    from ..parallel.embedding import ShardedEmbeddingCollection
    from ..parallel.embeddingbag import ShardedEmbeddingBagCollection

    kinds = (EmbeddingBagCollection, EmbeddingCollection, FusedEmbeddingBagCollection, FusedEmbeddingCollection, ShardedEmbeddingBagCollection, ShardedEmbeddingCollection)
    return [(n, m) for n, m in module.named_modules() if isinstance(m, kinds)]


def table_storages(collection: nn.Module) -> Dict[str, List[torch.Tensor]]:
    """table -> [weight rows, optimizer state rows...] of the rows THIS process holds (full table when unsharded)."""
    res: Dict[str, List[torch.Tensor]] = {}
    if hasattr(collection, "engine") and collection.engine is not None:
        for shard, w, st, _ in collection.engine.local_shard_views():
            assert shard.col_off == 0 and shard.cols == w.shape[1], "dynamic embedding needs row-complete shards (TW / RW)"
            res[shard.name] = [w] + [st[k] for k in sorted(st)]
        return res
    holder = getattr(collection, "embedding_bags", None) or getattr(collection, "embeddings")
    for name, m in holder.items():
        res[name] = [m.weight.data]
    return res


class DataLoaderIter:
    def __init__(self, it: Iterator, transform: Callable[[Any], Any], prefetch: int) -> None:
        self._q: "queue.Queue" = queue.Queue(maxsize=max(prefetch, 1))
        self._done = object()

        def run() -> None:
            try:
                for batch in it:
                    self._q.put(transform(batch))
            except BaseException as e:  # surface worker errors in the consumer
                self._q.put(e)
            self._q.put(self._done)

        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is self._done:
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        return item


class DataLoader:
    def __init__(self, collections: List[IDTransformerCollection], dataloader: Iterable, get_kjts: Callable[[Any], List[KeyedJaggedTensor]],
                 set_kjts: Callable[[Any, List[KeyedJaggedTensor]], Any], prefetch: int = 0) -> None:
        self._collections, self._dl, self._get, self._set, self._prefetch = collections, dataloader, get_kjts, set_kjts, prefetch

    def _transform(self, batch: Any) -> Any:
        kjts = self._get(batch)
        return self._set(batch, [c.transform(k) for c, k in zip(self._collections, kjts)])

    def __iter__(self):
        if self._prefetch <= 0:  # synchronous: transform (evict / fetch) strictly between two training steps
            return (self._transform(b) for b in self._dl)
        return DataLoaderIter(iter(self._dl), self._transform, self._prefetch)

    def __len__(self) -> int:
        return len(self._dl)  # type: ignore[arg-type]


def wrap(url: str, dataloader: Iterable, module: nn.Module, configs_dict: Optional[Dict[str, List]] = None, *, eviction_config: Optional[dict] = None,
         transform_config: Optional[dict] = None, get_kjts: Optional[Callable] = None, set_kjts: Optional[Callable] = None, prefetch: int = 0):
    """Attach a PS at ``url`` to every embedding collection of ``module`` and return ``(DataLoader, collections)``: the
    loader yields batches whose sparse ids were turned into cache ids, with the needed rows already resident.
    ``prefetch=0`` (default) transforms between steps. With ``prefetch>=1`` a worker thread transforms future batches
    while older ones still train: only safe when the cache is much larger than ``prefetch+1`` batches of ids, because a
    row evicted for batch i+k must not be in use by batch i."""
    colls: List[IDTransformerCollection] = []
    for path, m in _find_collections(module):
        cfgs = (configs_dict or {}).get(path) or (m.embedding_bag_configs() if hasattr(m, "embedding_bag_configs") else m.embedding_configs())
        stor = table_storages(m)
        ps = {}
        for c in cfgs:
            if c.name not in stor:
                continue  # table held by another rank
            tensors = stor[c.name]
            lo, hi = c.get_weight_init_min(), c.get_weight_init_max()

            def init(n, tensors=tensors, lo=lo, hi=hi):
                return [torch.empty(n, *tensors[0].shape[1:]).uniform_(lo, hi)] + [torch.zeros(n, *t.shape[1:]) for t in tensors[1:]]

            ps[c.name] = PS(f"{path}.{c.name}" if path else c.name, tensors, url, init_fn=init)
        colls.append(IDTransformerCollection([c for c in cfgs if c.name in stor], eviction_config, transform_config, ps))
    get_kjts = get_kjts or (lambda b: [b.sparse_features] * len(colls))
    if set_kjts is None:
        def set_kjts(b, kjts):
            b.sparse_features = kjts[0]
            return b
    return DataLoader(colls, dataloader, get_kjts, set_kjts, prefetch), colls
