"""Python face of the native id transformer (csrc/dynemb/id_map.{h,cpp}).
Parity: reference contrib/dynamic_embedding/src/torchrec_dynamic_embedding/id_transformer.py + `tde.IDTransformer`."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np
import torch

from ..csrc import build as _native

_STRATEGIES = {"mixed_lru_lfu": 0, "mixed_lfu_lru": 0, "lru": 1, "lfu": 2, "distance_lfu": 3}
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = _native.load("dynemb")
        i64p, u8p = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_uint8)
        L.trb_idt_create.restype = ctypes.c_void_p
        L.trb_idt_create.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.trb_idt_destroy.argtypes = [ctypes.c_void_p]
        L.trb_idt_transform.restype = ctypes.c_int64
        L.trb_idt_transform.argtypes = [ctypes.c_void_p, i64p, ctypes.c_int64, i64p, ctypes.c_int64, ctypes.c_int]
        L.trb_idt_evict.restype = ctypes.c_int64
        L.trb_idt_evict.argtypes = [ctypes.c_void_p, ctypes.c_int64, i64p]
        for fn in (L.trb_idt_size, L.trb_idt_pending_fetch):
            fn.restype = ctypes.c_int64
            fn.argtypes = [ctypes.c_void_p]
        L.trb_idt_take_fetch.restype = ctypes.c_int64
        L.trb_idt_take_fetch.argtypes = [ctypes.c_void_p, i64p, ctypes.c_int64]
        L.trb_idt_save.restype = ctypes.c_int64
        L.trb_idt_save.argtypes = [ctypes.c_void_p, i64p, ctypes.c_int64]
        L.trb_ps_create.restype = ctypes.c_void_p
        L.trb_ps_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        L.trb_ps_destroy.argtypes = [ctypes.c_void_p]
        L.trb_ps_push_async.restype = ctypes.c_int64
        L.trb_ps_push_async.argtypes = [ctypes.c_void_p, i64p, ctypes.c_int64, u8p, ctypes.c_int64]
        L.trb_ps_pull_async.restype = ctypes.c_int64
        L.trb_ps_pull_async.argtypes = [ctypes.c_void_p, i64p, ctypes.c_int64, u8p, ctypes.c_int64, u8p]
        L.trb_ps_wait.restype = ctypes.c_int
        L.trb_ps_wait.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.trb_ps_size.restype = ctypes.c_int64
        L.trb_ps_size.argtypes = [ctypes.c_void_p]
        L.trb_io_load_plugin.restype = ctypes.c_int
        L.trb_io_load_plugin.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        _LIB = L
    return _LIB


def _i64p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


class IDTransformer:
    """global id -> cache slot in ``[0, num_embedding)`` with LFU/LRU eviction bookkeeping.

    ``transform`` never evicts by itself: when the cache is full it reports ``success=False`` and slot -1 for the ids that
    did not fit; the caller evicts (``evict(n)`` returns the (global id, slot) pairs to write back to the PS) and retries."""

    def __init__(self, num_embedding: int, eviction_config: Optional[dict] = None, transform_config: Optional[dict] = None) -> None:
        ev = eviction_config or {"type": "mixed_lru_lfu"}
        tr = transform_config or {"type": "naive"}
        self._num_embedding = num_embedding
        self._threads = int(tr.get("threads", 4))
        self._h = lib().trb_idt_create(num_embedding, _STRATEGIES[ev.get("type", "mixed_lru_lfu")], int(ev.get("min_used_freq_power", 5)),
                                       int(tr.get("partitions", 8 if num_embedding >= 4096 else 1)))
        self._time = 0

    def __del__(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h and _LIB is not None:
            _LIB.trb_idt_destroy(h)

    def transform(self, global_ids: torch.Tensor, time: Optional[int] = None) -> Tuple[torch.Tensor, bool, torch.Tensor]:
        """Returns (cache ids, all resolved?, new (global id, cache id) pairs to fetch from the PS)."""
        if time is None:
            self._time += 1
            time = self._time
        g = np.ascontiguousarray(global_ids.detach().cpu().numpy().astype(np.int64, copy=False))
        out = np.empty_like(g)
        n_ok = lib().trb_idt_transform(self._h, _i64p(g), g.size, _i64p(out), int(time), self._threads)
        return torch.from_numpy(out).view(global_ids.shape), n_ok == g.size, self.take_fetch()

    def take_fetch(self) -> torch.Tensor:
        n = lib().trb_idt_pending_fetch(self._h)
        pairs = np.empty((max(n, 0), 2), dtype=np.int64)
        if n > 0:
            lib().trb_idt_take_fetch(self._h, _i64p(pairs), n)
        return torch.from_numpy(pairs)

    def evict(self, num_to_evict: int) -> torch.Tensor:
        pairs = np.empty((max(num_to_evict, 0), 2), dtype=np.int64)
        k = lib().trb_idt_evict(self._h, int(num_to_evict), _i64p(pairs)) if num_to_evict > 0 else 0
        return torch.from_numpy(pairs[:k].copy())

    def save(self) -> torch.Tensor:
        """All (global id, cache id, eviction record) triples."""
        n = len(self)
        out = np.empty((n, 3), dtype=np.int64)
        k = lib().trb_idt_save(self._h, _i64p(out), n)
        return torch.from_numpy(out[:k].copy())

    def __len__(self) -> int:
        return int(lib().trb_idt_size(self._h))

    @property
    def num_embedding(self) -> int:
        return self._num_embedding
