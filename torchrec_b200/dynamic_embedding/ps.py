"""Parameter-server client: moves embedding rows (+ optimizer state rows) between the cache table and an IO backend
(reference contrib/dynamic_embedding/src/torchrec_dynamic_embedding/ps.py, csrc/dynamic_embedding/ps.{h,cpp}).

A PS row is the concatenation of one row of every registered tensor (weight, momentum, ...) as raw bytes. Transfers are
staged through pinned host memory and run on the native IO thread pool, so ``evict``/``fetch`` overlap with training;
call ``wait()`` (or the next evict/fetch) to make them visible."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .id_transformer import _i64p, lib


def load_io_plugin(scheme: str, so_path: str) -> None:
    """Register an external IO backend (a ``.so`` exporting ``trb_io_plugin``) for ``<scheme>://`` urls."""
    if lib().trb_io_load_plugin(scheme.encode(), so_path.encode()) != 0:
        raise RuntimeError(f"cannot load IO plugin {so_path}")


class PS:
    def __init__(self, table_name: str, tensors: Sequence[torch.Tensor], url: str, init_fn=None, io_threads: int = 2, row_offset: int = 0) -> None:
        """tensors: per-row storages of the cache table, each ``[num_embedding, ...]`` (same first dim); row i of all tensors
        travels together. ``init_fn(n) -> List[Tensor]`` initialises rows the PS has never seen."""
        self.table_name = table_name
        self.tensors = list(tensors)
        self._row_elems = [int(t[0].numel()) if t.dim() > 1 else 1 for t in self.tensors]
        self._row_bytes_each = [e * t.element_size() for e, t in zip(self._row_elems, self.tensors)]
        self.row_bytes = sum(self._row_bytes_each)
        self._h = lib().trb_ps_create(table_name.encode(), url.encode(), io_threads)
        if not self._h:
            raise RuntimeError(f"cannot open IO backend {url}")
        self._init_fn = init_fn
        self._row_offset = row_offset
        self._inflight: List[Tuple[int, tuple]] = []  # (ticket, keep-alive buffers / completion)

    def __del__(self) -> None:
        try:
            self.wait()
        except Exception:
            pass
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().trb_ps_destroy(h)

    # ---- write back ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def evict(self, ids_to_evict: torch.Tensor) -> None:
        """ids_to_evict: ``[n, 2]`` (global id, cache id) pairs — their rows are pushed to the PS."""
        n = ids_to_evict.shape[0]
        if n == 0:
            return
        gids = np.ascontiguousarray(ids_to_evict[:, 0].numpy().astype(np.int64))
        slots = (ids_to_evict[:, 1] - self._row_offset).to(self.tensors[0].device)
        blob = torch.empty(n, self.row_bytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        c = 0
        for t, nb in zip(self.tensors, self._row_bytes_each):
            rows = t[slots].reshape(n, -1).contiguous().view(torch.uint8).reshape(n, nb)
            blob[:, c : c + nb].copy_(rows, non_blocking=False)
            c += nb
        ptr = ctypes.cast(blob.data_ptr(), ctypes.POINTER(ctypes.c_uint8))
        ticket = lib().trb_ps_push_async(self._h, _i64p(gids), n, ptr, self.row_bytes)
        self._inflight.append((ticket, (gids, blob, None)))

    # ---- read ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def fetch(self, ids_to_fetch: torch.Tensor, wait: bool = True) -> None:
        """ids_to_fetch: ``[n, 2]`` (global id, cache id) pairs — rows known to the PS are loaded into the cache table,
        unknown ones are initialised with ``init_fn``."""
        n = ids_to_fetch.shape[0]
        if n == 0:
            return
        self.wait()  # a pending eviction of the same id must land first
        gids = np.ascontiguousarray(ids_to_fetch[:, 0].numpy().astype(np.int64))
        slots = (ids_to_fetch[:, 1] - self._row_offset).clone()
        blob = torch.empty(n, self.row_bytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        found = np.zeros(n, dtype=np.uint8)
        ticket = lib().trb_ps_pull_async(self._h, _i64p(gids), n, ctypes.cast(blob.data_ptr(), ctypes.POINTER(ctypes.c_uint8)), self.row_bytes,
                                         found.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        self._inflight.append((ticket, (gids, blob, (slots, found))))
        if wait:
            self.wait()

    def wait(self) -> None:
        pending, self._inflight = self._inflight, []
        for ticket, (gids, blob, fetch_info) in pending:
            rc = lib().trb_ps_wait(self._h, ticket)
            if rc != 0:
                raise RuntimeError(f"PS IO failed for table {self.table_name} (rc={rc})")
            if fetch_info is not None:
                self._scatter(blob, *fetch_info)

    @torch.no_grad()
    def _scatter(self, blob: torch.Tensor, slots: torch.Tensor, found: np.ndarray) -> None:
        hit = torch.from_numpy(found.astype(bool))
        dev = self.tensors[0].device
        if bool(hit.any()):
            hs = slots[hit].to(dev)
            hb = blob[hit]
            c = 0
            for t, nb in zip(self.tensors, self._row_bytes_each):
                rows = hb[:, c : c + nb].contiguous().view(t.dtype).reshape(int(hit.sum()), *t.shape[1:])
                t[hs] = rows.to(dev, non_blocking=True)
                c += nb
        miss = ~hit
        if bool(miss.any()) and self._init_fn is not None:
            ms = slots[miss].to(dev)
            for t, v in zip(self.tensors, self._init_fn(int(miss.sum()))):
                t[ms] = v.to(dev).to(t.dtype)

    def __len__(self) -> int:
        return int(lib().trb_ps_size(self._h))
