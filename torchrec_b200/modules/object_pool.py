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






# ---- moved to ``tensor_pool.py`` (their reference import path); still importable from here ----
_MOVED_TO_TENSOR_POOL = ('TensorPool',)


def __getattr__(name: str):
    if name in _MOVED_TO_TENSOR_POOL:
        from . import tensor_pool as _m

        return getattr(_m, name)
    if name in ('KeyedJaggedTensorPool',):
        from . import keyed_jagged_tensor_pool as _m

        return getattr(_m, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
