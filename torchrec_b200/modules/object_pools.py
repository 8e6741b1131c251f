"""Object pools: id-addressed stores of dense rows (TensorPool) and jagged rows (KeyedJaggedTensorPool)
(reference torchrec/modules/tensor_pool.py:28, keyed_jagged_tensor_pool.py:77, object_pool_lookups.py)."""
from __future__ import annotations

import abc
from typing import Dict, Generic, List, Optional, Tuple, TypeVar

import torch
from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor

T = TypeVar("T")


class ObjectPool(abc.ABC, nn.Module, Generic[T]):
    """``lookup(ids) -> T`` / ``update(ids, values)``."""

    @abc.abstractmethod
    def lookup(self, ids: torch.Tensor) -> T:
        ...

    @abc.abstractmethod
    def update(self, ids: torch.Tensor, values: T) -> None:
        ...


class TensorPool(ObjectPool[torch.Tensor]):
    """``pool_size`` rows of ``dim`` values addressed by id (e.g. cached user embeddings)."""

    def __init__(self, pool_size: int, dim: int, dtype: torch.dtype, sharding_env=None, sharding_plan=None, device: Optional[torch.device] = None,
                 loading_required: bool = False, enable_uvm: bool = False) -> None:
        super().__init__()
        self._pool_size, self._dim, self._dtype = pool_size, dim, dtype
        self._device = device if device is not None else torch.device("cpu")
        self._enable_uvm = enable_uvm
        store_device = torch.device("cpu") if enable_uvm else self._device
        self.register_buffer("_pool", torch.zeros(pool_size, dim, dtype=dtype, device=store_device, pin_memory=enable_uvm and torch.cuda.is_available()))

    @property
    def pool_size(self) -> int:
        return self._pool_size

    @property
    def dim(self) -> int:
        return self._dim

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype

    @property
    def device(self) -> torch.device:
        return self._device

    @property
    def pool(self) -> torch.Tensor:
        return self._pool

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        out = self._pool[ids.to(self._pool.device).long()]
        return out.to(ids.device)

    def update(self, ids: torch.Tensor, values: torch.Tensor) -> None:
        assert values.dim() == 2 and values.size(1) == self._dim and values.dtype == self._dtype
        self._pool[ids.to(self._pool.device).long()] = values.to(self._pool.device)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return self.lookup(ids)

    def set_device(self, device_str: str) -> None:
        self._device = torch.device(device_str)


class KeyedJaggedTensorPool(ObjectPool[KeyedJaggedTensor]):
    """Per id, one jagged row (up to ``feature_max_lengths[key]`` values) per key."""

    def __init__(self, pool_size: int, feature_max_lengths: Dict[str, int], values_dtype: torch.dtype = torch.int64, is_weighted: bool = False,
                 device: Optional[torch.device] = None, enable_uvm: bool = False) -> None:
        super().__init__()
        self._pool_size = pool_size
        self._feature_max_lengths = feature_max_lengths
        self._keys = list(feature_max_lengths.keys())
        self._total = sum(feature_max_lengths.values())
        self._offsets = [0]
        for k in self._keys:
            self._offsets.append(self._offsets[-1] + feature_max_lengths[k])
        self._device = device if device is not None else torch.device("cpu")
        self._is_weighted = is_weighted
        self.register_buffer("_values", torch.zeros(pool_size, self._total, dtype=values_dtype, device=self._device))
        self.register_buffer("_lengths", torch.zeros(pool_size, len(self._keys), dtype=torch.int64, device=self._device))
        if is_weighted:
            self.register_buffer("_weights", torch.zeros(pool_size, self._total, dtype=torch.float32, device=self._device))

    @property
    def pool_size(self) -> int:
        return self._pool_size

    @property
    def feature_max_lengths(self) -> Dict[str, int]:
        return self._feature_max_lengths

    def lookup(self, ids: torch.Tensor) -> KeyedJaggedTensor:
        ids = ids.long()
        B = ids.numel()
        lengths = self._lengths[ids]  # [B, F]
        vals, ws = [], []
        for fi, k in enumerate(self._keys):
            block = self._values[ids, self._offsets[fi] : self._offsets[fi + 1]]
            mask = torch.arange(block.shape[1], device=block.device).unsqueeze(0) < lengths[:, fi : fi + 1]
            vals.append(block[mask])
            if self._is_weighted:
                ws.append(self._weights[ids, self._offsets[fi] : self._offsets[fi + 1]][mask])
        return KeyedJaggedTensor(keys=self._keys, values=torch.cat(vals), lengths=lengths.t().reshape(-1), weights=torch.cat(ws) if ws else None, stride=B)

    def update(self, ids: torch.Tensor, values: KeyedJaggedTensor) -> None:
        ids = ids.long()
        jt = values.to_dict()
        for fi, k in enumerate(self._keys):
            f = jt[k]
            mx = self._feature_max_lengths[k]
            dense = f.to_padded_dense(mx)
            self._values[ids, self._offsets[fi] : self._offsets[fi + 1]] = dense.to(self._values.dtype)
            self._lengths[ids, fi] = f.lengths().long().clamp(max=mx)
            if self._is_weighted and f.weights_or_none() is not None:
                self._weights[ids, self._offsets[fi] : self._offsets[fi + 1]] = f.to_padded_dense_weights(mx)

    def forward(self, ids: torch.Tensor) -> KeyedJaggedTensor:
        return self.lookup(ids)
