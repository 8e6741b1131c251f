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

    # ---- tensor protocol: the ops state-dict / DTensor / checkpoint plumbing applies to a state value ------------------------------------------
    # Per-shard ops (applied to every local shard, result re-wrapped with the same offsets), in-place per-shard ops, and a few
    # whole-object ops. Arithmetic is intentionally absent: this is a container, not a math tensor.
    _PER_SHARD = ("detach", "clone", "_to_copy", "contiguous", "alias", "zeros_like", "empty_like", "ones_like", "_pin_memory", "lift_fresh")
    _PER_SHARD_INPLACE = ("zero_", "fill_", "requires_grad_", "record_stream")

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):  # type: ignore[override]
        kwargs = kwargs or {}
        name = getattr(getattr(func, "overloadpacket", func), "__name__", None) or getattr(func, "__name__", str(func))
        self = next(a for a in list(args) + list(kwargs.values()) if isinstance(a, LocalShardsWrapper))
        rest = args[1:]
        if name in cls._PER_SHARD:
            return LocalShardsWrapper([func(s, *rest, **kwargs) for s in self._local_shards], self._offsets)
        if name in cls._PER_SHARD_INPLACE:
            for s in self._local_shards:
                func(s, *rest, **kwargs)
            return self
        if name in ("view", "_unsafe_view", "reshape"):
            # DTensor.from_local takes `input.view_as(input)`: the only views that make sense keep the bounding shape
            shape = tuple(int(x) for x in rest[0])
            if shape != tuple(self.shape):
                raise NotImplementedError(f"LocalShardsWrapper.view to {shape}: only the identity view of {tuple(self.shape)} is defined")
            return LocalShardsWrapper(list(self._local_shards), self._offsets)
        if name == "copy_":
            src = args[1]
            if isinstance(src, LocalShardsWrapper):
                if len(src._local_shards) != len(self._local_shards):
                    raise ValueError(f"copy_ between wrappers with {len(self._local_shards)} and {len(src._local_shards)} local shards")
                for d, s_ in zip(self._local_shards, src._local_shards):
                    d.copy_(s_)
            else:  # a full tensor: take every shard's window
                for d, off in zip(self._local_shards, self._offsets):
                    d.copy_(src[tuple(slice(o, o + n) for o, n in zip(off, d.shape))])
            return self
        if name == "equal":
            other = args[1] if args[0] is self else args[0]
            return (isinstance(other, LocalShardsWrapper) and self._offsets == other._offsets
                    and all(torch.equal(x, y) for x, y in zip(self._local_shards, other._local_shards)))
        if name == "is_pinned":
            return all(s.is_pinned() for s in self._local_shards)
        if name in ("sym_size", "sym_stride", "sym_numel", "sym_storage_offset", "dim", "is_contiguous"):
            return func(torch.empty(self.shape, device="meta"), *rest, **kwargs)
        raise NotImplementedError(f"LocalShardsWrapper does not implement {name}: use .local_shards()")

    # ---- torch.compile / tracing --------------------------------------------------------------------------------------------------------------
    def __tensor_flatten__(self):
        return [f"_shard_{i}" for i in range(len(self._local_shards))], {"offsets": self._offsets}

    @staticmethod
    def __tensor_unflatten__(inner_tensors, meta, outer_size=None, outer_stride=None):
        n = len(inner_tensors)
        return LocalShardsWrapper([inner_tensors[f"_shard_{i}"] for i in range(n)], meta["offsets"])

    def __getattr__(self, item):  # `_shard_<i>` attributes named by __tensor_flatten__
        if item.startswith("_shard_") and item[7:].isdigit():
            return self._local_shards[int(item[7:])]
        raise AttributeError(item)

    # ---- torch.distributed.checkpoint hooks (one write item / chunk per local shard, addressed by its global offsets) --------------------------
    def __create_chunk_list__(self):
        from torch.distributed.checkpoint.metadata import ChunkStorageMetadata

        return [ChunkStorageMetadata(offsets=torch.Size(o), sizes=torch.Size(s.shape)) for s, o in zip(self._local_shards, self._offsets)]

    def __create_write_items__(self, fqn: str, object: Any):
        from torch.distributed.checkpoint.metadata import ChunkStorageMetadata, MetadataIndex, TensorProperties
        from torch.distributed.checkpoint.planner import TensorWriteData, WriteItem, WriteItemType

        size = object.size() if hasattr(object, "size") else self.size()
        return [WriteItem(index=MetadataIndex(fqn, torch.Size(o)), type=WriteItemType.SHARD,
                          tensor_data=TensorWriteData(chunk=ChunkStorageMetadata(offsets=torch.Size(o), sizes=torch.Size(s.shape)),
                                                      properties=TensorProperties.create_from_tensor(s), size=size))
                for s, o in zip(self._local_shards, self._offsets)]

    def __get_tensor_shard__(self, index) -> torch.Tensor:
        """Local shard addressed by a checkpoint ``MetadataIndex`` (its ``offset``, optionally the positional hint ``index``)."""
        if index.offset is not None:
            hint = getattr(index, "index", None)
            if hint is not None and hint < len(self._offsets) and tuple(self._offsets[hint]) == tuple(index.offset):
                return self._local_shards[hint]
            for s, o in zip(self._local_shards, self._offsets):
                if tuple(o) == tuple(index.offset):
                    return s
        raise ValueError(f"no local shard at offset {index.offset} (have {self._offsets})")

    def _get_tensor_size_bytes(self) -> int:
        return sum(s.numel() * s.element_size() for s in self._local_shards)

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
