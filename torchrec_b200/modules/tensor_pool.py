"""``TensorPool``: a table of fixed-width tensors addressed by id (lookup / update), shardable row-wise (reference ``torchrec/modules/tensor_pool.py:28``)."""
from __future__ import annotations

import abc
from typing import Dict, Generic, List, Optional, Tuple, TypeVar
import torch
from torch import nn
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .object_pool import ObjectPool  # noqa: F401


class TensorPool(ObjectPool[torch.Tensor]):
    """``pool_size`` rows of ``dim`` values addressed by id (e.g. cached user embeddings)."""

    def __init__(self, pool_size: int, dim: int, dtype: torch.dtype, sharding_env=None, sharding_plan=None, device: Optional[torch.device] = None,
                 loading_required: bool = False, enable_uvm: bool = False, loaded_values: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self._pool_size, self._dim, self._dtype = pool_size, dim, dtype
        self._device = device if device is not None else torch.device("cpu")
        self._enable_uvm = enable_uvm
        store_device = torch.device("cpu") if enable_uvm else self._device
        self.register_buffer("_pool", torch.zeros(pool_size, dim, dtype=dtype, device=store_device, pin_memory=enable_uvm and torch.cuda.is_available()))
        if loaded_values is not None:  # pre-computed rows instead of zeros
            assert tuple(loaded_values.shape) == (pool_size, dim), f"loaded_values must be [{pool_size}, {dim}], got {tuple(loaded_values.shape)}"
            self._pool.copy_(loaded_values.to(dtype))

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
