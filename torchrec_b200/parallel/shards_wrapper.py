"""``LocalShardsWrapper``: the local shards of one logical tensor + their offsets, as ONE tensor-like object (reference
torchrec/distributed/shards_wrapper.py:30-392).

It is what a DTensor-style state dict carries for tables that own several shards on a rank (column-wise: several column blocks of one
table). Implemented as a ``torch.Tensor`` wrapper subclass so it travels through ``state_dict`` / ``torch.save`` / ``.to`` / ``detach``
/ ``clone`` / ``zero_``-style calls; arithmetic is intentionally not supported (checkpoint plumbing only)."""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import torch


class LocalShardsWrapper(torch.Tensor):
    __slots__ = ["_local_shards", "_offsets", "_sizes"]

    @staticmethod
    def __new__(cls, local_shards: List[torch.Tensor], local_offsets: List[Tuple[int, ...]]) -> "LocalShardsWrapper":
        assert len(local_shards) == len(local_offsets)
        if local_shards:
            first = local_shards[0]
            assert all(s.dim() == first.dim() for s in local_shards)
            # bounding size: concatenation along the last dim when rows agree (column-wise), else along dim 0
            if first.dim() == 2 and all(s.shape[0] == first.shape[0] for s in local_shards):
                size = (first.shape[0], sum(s.shape[1] for s in local_shards))
            else:
                size = (sum(s.shape[0] for s in local_shards),) + tuple(first.shape[1:])
            dtype, device, requires_grad = first.dtype, first.device, first.requires_grad
        else:
            size, dtype, device, requires_grad = (0,), torch.float32, torch.device("cpu"), False
        r = torch.Tensor._make_wrapper_subclass(cls, size, dtype=dtype, device=device, requires_grad=requires_grad)  # type: ignore[attr-defined]
        r._local_shards = list(local_shards)
        r._offsets = [tuple(int(x) for x in o) for o in local_offsets]
        r._sizes = [tuple(s.shape) for s in local_shards]
        return r

    # ---- accessors ---------------------------------------------------------------------------------------------------------------
    def local_shards(self) -> List[torch.Tensor]:
        return self._local_shards

    def local_offsets(self) -> List[Tuple[int, ...]]:
        return self._offsets

    def local_sizes(self) -> List[Tuple[int, ...]]:
        return self._sizes

    def shards_with_offsets(self) -> List[Tuple[torch.Tensor, Tuple[int, ...]]]:
        return list(zip(self._local_shards, self._offsets))

    def __repr__(self) -> str:  # type: ignore[override]
        return f"LocalShardsWrapper(shards={[tuple(s.shape) for s in self._local_shards]}, offsets={self._offsets})"

    # ---- the handful of ops checkpoint code applies to state-dict values -----------------------------------------------------------------
    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):  # type: ignore[override]
        kwargs = kwargs or {}
        name = func.__name__ if hasattr(func, "__name__") else str(func)
        self = next(a for a in list(args) + list(kwargs.values()) if isinstance(a, LocalShardsWrapper))
        unary = ("detach", "clone", "_to_copy", "zero_", "contiguous", "alias", "view_as_real")
        if any(name.startswith(u) for u in unary):
            rest = args[1:]
            shards = [func(s, *rest, **kwargs) for s in self._local_shards]
            if name.startswith("zero_"):
                return self
            return LocalShardsWrapper(shards, self._offsets)
        if name.startswith("copy_"):
            src = args[1]
            if isinstance(src, LocalShardsWrapper):
                for d, s in zip(self._local_shards, src._local_shards):
                    d.copy_(s)
            else:  # a full tensor: take every shard's window
                for d, off in zip(self._local_shards, self._offsets):
                    idx = tuple(slice(o, o + n) for o, n in zip(off, d.shape))
                    d.copy_(src[idx])
            return self
        if name.startswith("equal"):
            other = args[1] if args[0] is self else args[0]
            return isinstance(other, LocalShardsWrapper) and self._offsets == other._offsets and all(torch.equal(a, b) for a, b in zip(self._local_shards, other._local_shards))
        raise NotImplementedError(f"LocalShardsWrapper does not implement {name}: use .local_shards()")

    def __reduce_ex__(self, protocol):  # picklable for torch.save
        return (LocalShardsWrapper, (self._local_shards, self._offsets))

    def full_tensor(self, global_size: Optional[Tuple[int, ...]] = None) -> torch.Tensor:
        """Dense tensor with the local shards placed at their offsets (zeros elsewhere); ``global_size`` defaults to the bounding box."""
        if not self._local_shards:
            return torch.empty(0)
        if global_size is None:
            global_size = tuple(max(o[d] + s.shape[d] for s, o in zip(self._local_shards, self._offsets)) for d in range(self._local_shards[0].dim()))
        out = torch.zeros(global_size, dtype=self.dtype, device=self._local_shards[0].device)
        for s, off in zip(self._local_shards, self._offsets):
            out[tuple(slice(o, o + n) for o, n in zip(off, s.shape))] = s
        return out
