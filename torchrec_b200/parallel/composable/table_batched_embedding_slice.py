"""``nn.Parameter`` view over one table's slice of a flat table-batched weight buffer (reference
torchrec/distributed/composable/table_batched_embedding_slice.py:16-107).

The engine stores all local shards of a kernel group in ONE flat buffer (one launch looks up all of them); optimizers, DDP-ignore
lists and ``named_parameters`` want one parameter per table. The slice shares storage with the buffer: in-place kernel updates are
visible through it, ``.data`` assignment of the buffer must be followed by ``rebind``."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn


class TableBatchedEmbeddingSlice(nn.Parameter):
    __slots__ = ["_original_tensor", "_start_offset", "_end_offset", "_num_embeddings", "_embedding_dim"]

    def __new__(cls, data: torch.Tensor, start_offset: int, end_offset: int, num_embeddings: int, embedding_dim: int) -> "TableBatchedEmbeddingSlice":
        flat = data.detach().reshape(-1)
        view = flat[start_offset:end_offset].view(num_embeddings, embedding_dim)
        obj = nn.Parameter.__new__(cls, view, requires_grad=data.requires_grad)  # type: ignore[call-arg]
        obj._original_tensor = data
        obj._start_offset, obj._end_offset = int(start_offset), int(end_offset)
        obj._num_embeddings, obj._embedding_dim = int(num_embeddings), int(embedding_dim)
        return obj

    def __init__(self, *args, **kwargs) -> None:  # nn.Parameter is built in __new__
        pass

    def __deepcopy__(self, memo):
        if id(self) in memo:
            return memo[id(self)]
        data = self._original_tensor.detach().clone()
        out = TableBatchedEmbeddingSlice(data.requires_grad_(self.requires_grad), self._start_offset, self._end_offset, self._num_embeddings, self._embedding_dim)
        memo[id(self)] = out
        return out

    def rebind(self, data: torch.Tensor) -> "TableBatchedEmbeddingSlice":
        """New slice over ``data`` (same offsets) after the flat buffer was re-allocated (resharding, FULLY_SHARDED gather)."""
        return TableBatchedEmbeddingSlice(data, self._start_offset, self._end_offset, self._num_embeddings, self._embedding_dim)

    @property
    def grad(self) -> Optional[torch.Tensor]:
        g = self._original_tensor.grad
        if g is None:
            return None
        return g.reshape(-1)[self._start_offset : self._end_offset].view(self._num_embeddings, self._embedding_dim)

    @grad.setter
    def grad(self, value: Optional[torch.Tensor]) -> None:
        if value is None:
            return
        if self._original_tensor.grad is None:
            self._original_tensor.grad = torch.zeros_like(self._original_tensor)
        self._original_tensor.grad.reshape(-1)[self._start_offset : self._end_offset].copy_(value.reshape(-1))

    @property
    def grad_fn(self) -> None:  # a leaf for autograd purposes
        return None
