"""Local storage + lookup/update of object-pool shards.

Reference: ``torchrec/modules/object_pool_lookups.py`` - ``KeyedJaggedTensorPoolLookup`` :28, ``TensorJaggedIndexSelectLookup`` :156, ``UVMCachingInt64Lookup`` :298,
``UVMCachingInt32Lookup`` :440, ``TensorPoolLookup`` :572, ``TensorLookup`` :637, ``UVMCachingFloatLookup`` :695. The reference backs its "UVM caching" lookups
with an FBGEMM TBE used as a key-value store; here they sit on ``UvmCachedEmbeddingBags`` (host rows + HBM row cache, ``ops/uvm.py``) on a GPU and on
a plain tensor on CPU. Int64 values are stored as two fp32-sized words per element (bit-exact ``view``), int32 as one.
"""
from __future__ import annotations

import abc
from typing import Dict, Iterator, List, Optional, Tuple

import torch

from ..sparse.jagged_tensor import JaggedTensor


def _dense_to_jagged_rows(rows: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """Keep the first ``lengths[i]`` entries of every padded row."""
    mask = torch.arange(rows.shape[1], device=rows.device).unsqueeze(0) < lengths.unsqueeze(1)
    return rows[mask]


class KeyedJaggedTensorPoolLookup(abc.ABC, torch.nn.Module):
    """Pool of KJT rows: per id, ``F`` features of at most ``feature_max_lengths[f]`` values. Storage = padded values ``[pool, sum(max_len)]`` +
    per-(id, feature) lengths; ``lookup`` returns a JaggedTensor whose lengths are ordered (feature, id) like a KJT's."""

    def __init__(self, pool_size: int, feature_max_lengths: Dict[str, int], is_weighted: bool, device: torch.device) -> None:
        super().__init__()
        self._pool_size = pool_size
        self._feature_max_lengths = feature_max_lengths
        self._device = device
        self._total_lengths = sum(feature_max_lengths.values())
        self._is_weighted = is_weighted
        self._key_lengths = torch.zeros(pool_size, len(feature_max_lengths), dtype=torch.int32, device=device)
        offs = [0]
        for v in feature_max_lengths.values():
            offs.append(offs[-1] + v)
        self._feature_offsets = offs

    @abc.abstractmethod
    def _read(self, ids: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        ...

    @abc.abstractmethod
    def _write(self, ids: torch.Tensor, values: torch.Tensor, weights: Optional[torch.Tensor]) -> None:
        ...

    def lookup(self, ids: torch.Tensor) -> JaggedTensor:
        ids = ids.long()
        rows, wrows = self._read(ids)
        lens = self._key_lengths[ids]  # [n, F]
        vals, ws, out_lens = [], [], []
        for f in range(lens.shape[1]):
            a, b = self._feature_offsets[f], self._feature_offsets[f + 1]
            vals.append(_dense_to_jagged_rows(rows[:, a:b], lens[:, f].long()))
            if wrows is not None:
                ws.append(_dense_to_jagged_rows(wrows[:, a:b], lens[:, f].long()))
            out_lens.append(lens[:, f])
        return JaggedTensor(values=torch.cat(vals) if vals else rows.new_empty(0), weights=torch.cat(ws) if ws else None, lengths=torch.cat(out_lens) if out_lens else lens.new_empty(0))

    def update(self, ids: torch.Tensor, values: JaggedTensor) -> None:
        ids = ids.long()
        n, F = ids.numel(), len(self._feature_max_lengths)
        lens = values.lengths().view(F, n).t().contiguous()  # [n, F]
        maxes = torch.tensor(list(self._feature_max_lengths.values()), device=lens.device)
        lens = torch.minimum(lens.long(), maxes)
        dense = torch.zeros(n, self._total_lengths, dtype=values.values().dtype, device=values.values().device)
        wdense = torch.zeros(n, self._total_lengths, dtype=torch.float32, device=dense.device) if self._is_weighted else None
        src_off = torch.cumsum(values.lengths().long(), 0) - values.lengths().long()
        for f in range(F):
            a = self._feature_offsets[f]
            for i in range(n):
                L = int(lens[i, f])
                if L:
                    s = int(src_off[f * n + i])
                    dense[i, a : a + L] = values.values()[s : s + L]
                    if wdense is not None and values.weights_or_none() is not None:
                        wdense[i, a : a + L] = values.weights()[s : s + L]
        self._key_lengths[ids] = lens.to(self._key_lengths.dtype).to(self._key_lengths.device)
        self._write(ids, dense, wdense)

    def forward(self, ids: torch.Tensor) -> JaggedTensor:
        return self.lookup(ids)

    def states_to_register(self) -> Iterator[Tuple[str, torch.Tensor]]:
        yield "key_lengths", self._key_lengths


class TensorJaggedIndexSelectLookup(KeyedJaggedTensorPoolLookup):
    """Values in a plain device tensor (HBM)."""

    def __init__(self, pool_size: int, values_dtype: torch.dtype, feature_max_lengths: Dict[str, int], is_weighted: bool, device: torch.device) -> None:
        super().__init__(pool_size, feature_max_lengths, is_weighted, device)
        self._values = torch.zeros(pool_size, self._total_lengths, dtype=values_dtype, device=device)
        self._weights = torch.zeros(pool_size, self._total_lengths, dtype=torch.float32, device=device) if is_weighted else None

    def _read(self, ids):
        return self._values[ids], (self._weights[ids] if self._weights is not None else None)

    def _write(self, ids, values, weights):
        self._values[ids] = values.to(self._values.dtype).to(self._values.device)
        if self._weights is not None and weights is not None:
            self._weights[ids] = weights.to(self._weights.device)

    def states_to_register(self):
        yield "values", self._values
        yield "key_lengths", self._key_lengths
        if self._weights is not None:
            yield "weights", self._weights


class _CachedRows(torch.nn.Module):
    """``[rows, words]`` fp32 words in host memory behind an HBM row cache on a GPU; a device tensor elsewhere."""

    def __init__(self, rows: int, words: int, device: torch.device) -> None:
        super().__init__()
        self._words = words
        self._padded = -(-words // 4) * 4
        self._tbe = None
        if device.type == "cuda":
            from ..ops.tbe import OptimType, PoolingMode
            from ..ops.uvm import UvmCachedEmbeddingBags

            self._tbe = UvmCachedEmbeddingBags([(rows, self._padded)], [0], cache_load_factor=0.2, device=device, pooling_mode=PoolingMode.NONE, optimizer=OptimType.EXACT_SGD,
                                               learning_rate=0.0)
            with torch.no_grad():
                self._tbe.split_embedding_weights()[0].zero_()
                self._tbe.load_rows_changed()
        else:
            self._store = torch.zeros(rows, self._padded, dtype=torch.float32, device=device)

    def read(self, ids: torch.Tensor) -> torch.Tensor:
        if self._tbe is None:
            return self._store[ids][:, : self._words]
        offsets = torch.arange(ids.numel() + 1, device=ids.device, dtype=torch.int64)
        with torch.no_grad():
            return self._tbe(ids, offsets, batch_size=ids.numel())[:, : self._words]

    @torch.no_grad()
    def write(self, ids: torch.Tensor, rows: torch.Tensor) -> None:
        if self._tbe is None:
            self._store[ids, : self._words] = rows
            return
        self._tbe.flush()
        self._tbe.split_embedding_weights(flush=False)[0][ids.cpu() if not self._tbe.split_embedding_weights(flush=False)[0].is_cuda else ids, : self._words] = rows.to(
            self._tbe.split_embedding_weights(flush=False)[0].device)
        self._tbe.load_rows_changed()

    def table(self) -> torch.Tensor:
        return self._store if self._tbe is None else self._tbe.split_embedding_weights()[0]


class UVMCachingInt64Lookup(KeyedJaggedTensorPoolLookup):
    """int64 values, two 32-bit words each, in cached host rows."""

    _WORDS = 2
    _DTYPE = torch.int64

    def __init__(self, pool_size: int, feature_max_lengths: Dict[str, int], is_weighted: bool, device: torch.device) -> None:
        super().__init__(pool_size, feature_max_lengths, is_weighted, device)
        self._rows = _CachedRows(pool_size, self._total_lengths * self._WORDS, device)
        self._wrows = _CachedRows(pool_size, self._total_lengths, device) if is_weighted else None

    def _read(self, ids):
        raw = self._rows.read(ids).contiguous()
        vals = raw.view(torch.int32).view(ids.numel(), -1)
        vals = vals.contiguous().view(self._DTYPE).view(ids.numel(), self._total_lengths) if self._WORDS == 2 else vals
        return vals, (self._wrows.read(ids) if self._wrows is not None else None)

    def _write(self, ids, values, weights):
        v = values.to(self._DTYPE).contiguous()
        self._rows.write(ids, v.view(torch.float32).view(ids.numel(), -1))
        if self._wrows is not None and weights is not None:
            self._wrows.write(ids, weights)

    def states_to_register(self):
        yield "values_upper_and_lower_bits", self._rows.table()
        yield "key_lengths", self._key_lengths


class UVMCachingInt32Lookup(UVMCachingInt64Lookup):
    _WORDS = 1
    _DTYPE = torch.int32

    def states_to_register(self):
        yield "values", self._rows.table()
        yield "key_lengths", self._key_lengths


# ---- tensor pools ----------------------------------------------------------------------------------------------------
class TensorPoolLookup(abc.ABC, torch.nn.Module):
    def __init__(self, pool_size: int, dim: int, dtype: torch.dtype, device: torch.device) -> None:
        super().__init__()
        self._pool_size, self._dim, self._dtype, self._device = pool_size, dim, dtype, device

    @abc.abstractmethod
    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        ...

    @abc.abstractmethod
    def update(self, ids: torch.Tensor, values: torch.Tensor) -> None:
        ...

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return self.lookup(ids)

    @abc.abstractmethod
    def states_to_register(self) -> Iterator[Tuple[str, torch.Tensor]]:
        ...

    @abc.abstractmethod
    def set_state(self, loaded_values: torch.Tensor) -> None:
        ...


class TensorLookup(TensorPoolLookup):
    def __init__(self, pool_size: int, dim: int, dtype: torch.dtype, device: torch.device, enable_uvm: bool = False) -> None:
        super().__init__(pool_size, dim, dtype, device)
        self._enable_uvm = enable_uvm
        if enable_uvm and device.type == "cuda":
            from ..ops.uvm import new_unified_tensor

            self._pool = new_unified_tensor(torch.zeros(1, dtype=dtype, device=device), [pool_size, dim], is_host_mapped=True)
            self._pool.zero_()
        else:
            self._pool = torch.zeros(pool_size, dim, dtype=dtype, device=device)

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        return self._pool[ids.long()]

    def update(self, ids: torch.Tensor, values: torch.Tensor) -> None:
        self._pool[ids.long()] = values.to(self._pool.dtype)

    def set_state(self, loaded_values: torch.Tensor) -> None:
        self._pool.copy_(loaded_values)

    def states_to_register(self):
        yield "_pool", self._pool


class UVMCachingFloatLookup(TensorPoolLookup):
    def __init__(self, pool_size: int, dim: int, dtype: torch.dtype, device: torch.device) -> None:
        super().__init__(pool_size, dim, dtype, device)
        self._rows = _CachedRows(pool_size, dim, device)

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        return self._rows.read(ids.long()).to(self._dtype)

    def update(self, ids: torch.Tensor, values: torch.Tensor) -> None:
        self._rows.write(ids.long(), values.float())

    def states_to_register(self):
        yield "_pool", self._rows.table()[:, : self._dim]

    def set_state(self, loaded_values: torch.Tensor) -> None:
        self._rows.write(torch.arange(self._pool_size, device=loaded_values.device), loaded_values.float())
