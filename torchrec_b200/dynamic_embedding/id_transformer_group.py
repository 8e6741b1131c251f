"""Id transformers of ALL dynamic-embedding collections of a model, each running on its own thread.

Reference: ``contrib/dynamic_embedding/.../id_transformer_group.py:38`` (``IDTransformerGroup``). A model may hold several embedding collections (say user
and item towers); ``transform`` takes ``{module path: KJT}``, runs every collection's global-id -> cache-id translation (with its evictions / fetches
through the parameter server) concurrently and returns ``({module path: cache-id KJT}, {module path: fetch handles})``. The translation is native code
that releases the GIL (``csrc/dynemb``), so the threads really overlap.
"""
from __future__ import annotations

import queue
import threading
from typing import Any, Dict, List, Optional, Tuple

from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor
from .id_transformer_collection import IDTransformerCollection
from .ps import PS


def _create_transformer_thread(transformer: IDTransformerCollection) -> Tuple[threading.Thread, "queue.Queue[Any]", "queue.Queue[Any]"]:
    """A daemon worker: pops KJTs from its input queue, pushes (transformed KJT | exception) to its output queue; ``None`` stops it."""
    inq: "queue.Queue[Any]" = queue.Queue()
    outq: "queue.Queue[Any]" = queue.Queue()

    def loop() -> None:
        while True:
            item = inq.get()
            if item is None:
                break
            try:
                outq.put(transformer.transform(item))
            except BaseException as e:  # surfaced in the caller's thread
                outq.put(e)

    t = threading.Thread(target=loop, daemon=True, name="trb-id-transformer")
    t.start()
    return t, inq, outq


class IDTransformerGroup:
    def __init__(self, url: str, module: nn.Module, configs_dict: Dict[str, List[Any]], *, eviction_config: Optional[dict] = None, transform_config: Optional[dict] = None,
                 ps_config: Optional[dict] = None, parallel: bool = True) -> None:
        """``configs_dict``: ``{path of an embedding collection inside module: its table configs}`` (``num_embeddings`` = cache rows).
        ``url``: parameter-server address (``memory://``, ``file://dir`` or a registered IO plugin scheme)."""
        from .dataloader import table_storages

        self._parallel = parallel
        self._collections: Dict[str, IDTransformerCollection] = {}
        self._threads: Dict[str, Tuple[threading.Thread, "queue.Queue[Any]", "queue.Queue[Any]"]] = {}
        for path, configs in configs_dict.items():
            collection = module.get_submodule(path) if path else module
            storages = table_storages(collection)
            ps = {}
            for cfg in configs:
                tensors = storages.get(cfg.name)
                if tensors is None:
                    raise KeyError(f"{path}: no storage found for table {cfg.name}")
                ps[cfg.name] = PS(f"{path}.{cfg.name}" if path else cfg.name, tensors, url, **(ps_config or {}))
            self._collections[path] = IDTransformerCollection(configs, eviction_config, transform_config, ps)
            if parallel:
                self._threads[path] = _create_transformer_thread(self._collections[path])

    def transform(self, kjt_dict: Dict[str, KeyedJaggedTensor]) -> Tuple[Dict[str, KeyedJaggedTensor], Dict[str, Any]]:
        """Translate every collection's batch; returns cache-id KJTs and, per path, the PS handles whose fetches the caller may still ``wait()`` on."""
        unknown = set(kjt_dict) - set(self._collections)
        if unknown:
            raise KeyError(f"no dynamic-embedding collection at {sorted(unknown)}; known: {sorted(self._collections)}")
        out: Dict[str, KeyedJaggedTensor] = {}
        if self._parallel:
            for path, kjt in kjt_dict.items():
                self._threads[path][1].put(kjt)
            for path in kjt_dict:
                res = self._threads[path][2].get()
                if isinstance(res, BaseException):
                    raise res
                out[path] = res
        else:
            for path, kjt in kjt_dict.items():
                out[path] = self._collections[path].transform(kjt)
        return out, {path: self._collections[path]._ps for path in kjt_dict}

    def save(self) -> None:
        """Write every cached row back to the parameter server (checkpoint)."""
        for c in self._collections.values():
            c.save()

    def __contains__(self, path: str) -> bool:
        return path in self._collections

    def __getitem__(self, path: str) -> IDTransformerCollection:
        return self._collections[path]

    def close(self) -> None:
        for t, inq, _ in self._threads.values():
            inq.put(None)
        for t, _, _ in self._threads.values():
            t.join(timeout=5)
        self._threads = {}

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass
