"""``KeyedJaggedTensorPool``: per-id jagged feature values with fixed per-key capacity (reference ``torchrec/modules/keyed_jagged_tensor_pool.py:77``)."""
from __future__ import annotations

import abc
from typing import Dict, Generic, List, Optional, Tuple, TypeVar
import torch
from torch import nn
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .object_pool import ObjectPool  # noqa: F401


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
