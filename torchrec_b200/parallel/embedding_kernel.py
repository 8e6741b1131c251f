"""Base class of the per-group embedding kernels + state-dict helpers.

Reference: ``torchrec/distributed/embedding_kernel.py`` (``BaseEmbedding`` :63-178, virtual-table metadata helpers :181-347, ``get_state_dict``
:350-447). A *kernel* here is one ``TableBatchedEmbeddingBags`` (``ops/tbe.py``: one flat weight buffer, single-launch lookup, fused optimizer) built
from one ``GroupedEmbeddingConfig``; the sharded modules normally reach the kernels through ``ShardedLookupEngine`` (``parallel/engine.py``), the
classes in ``batched_embedding_kernel.py`` expose the same kernels behind the reference's composable per-group API.
"""
from __future__ import annotations

import abc
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Dict, Iterator, List, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor
from .embedding_types import GroupedEmbeddingConfig, ShardedEmbeddingTable
from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardedTensorMetadata
from torch.distributed._shard.sharding_spec import ShardMetadata


@dataclass
class RawIdTrackerWrapper:
    """Callbacks a kernel uses to hand raw (pre-remap) ids to the model-delta tracker (reference :45-60)."""

    get_indexed_lookups: Any = None
    delete: Any = None


class BaseEmbedding(abc.ABC, nn.Module):
    """One lookup kernel over the tables of a ``GroupedEmbeddingConfig`` (reference :63)."""

    def __init__(self) -> None:
        super().__init__()
        self._raw_id_tracker_wrapper: Optional[RawIdTrackerWrapper] = None

    @abc.abstractmethod
    def forward(self, features: KeyedJaggedTensor) -> torch.Tensor:
        ...

    @property
    @abc.abstractmethod
    def config(self) -> GroupedEmbeddingConfig:
        ...

    def init_raw_id_tracker(self, get_indexed_lookups: Any, delete: Any) -> None:
        self._raw_id_tracker_wrapper = RawIdTrackerWrapper(get_indexed_lookups, delete)

    def _track_raw_ids(self, features: KeyedJaggedTensor) -> None:
        w = self._raw_id_tracker_wrapper
        if w is not None and w.get_indexed_lookups is not None:
            w.get_indexed_lookups(self.config.table_names(), features)

    def prefetch(self, features: KeyedJaggedTensor, forward_stream: Optional[torch.cuda.Stream] = None) -> None:
        """HBM-cached kernels stage the rows of ``features``; others have nothing to do."""

    def flush(self) -> None:
        pass

    def purge(self) -> None:
        pass


def _make_sharded_tensor(local_shards: List[Shard], global_metadata: Any, pg: Optional[dist.ProcessGroup]) -> Any:
    """A torch ShardedTensor when a process group exists, else the local shard(s) - no communication either way."""
    if pg is None or not dist.is_initialized() or not isinstance(global_metadata, ShardedTensorMetadata):
        return local_shards[0].tensor if len(local_shards) == 1 else torch.cat([s.tensor for s in local_shards], dim=1)
    return ShardedTensor._init_from_local_shards_and_global_metadata(local_shards, global_metadata, process_group=pg)


# ---- virtual (key-value) tables: a shard's "rows" are the ids currently materialised -------------------------------
def create_virtual_table_local_metadata(local_metadata: ShardMetadata, param: Union[torch.Tensor, Any], my_rank: int,
                                        offset: Optional[int] = None, weight_count_per_rank: Optional[List[int]] = None) -> None:
    """Rewrite ``local_metadata`` in place so it describes the rows that currently exist in the KV shard (reference :181-201)."""
    rows = int(param.size(0)) if hasattr(param, "size") else int(param)
    if offset is None:
        offset = my_rank if weight_count_per_rank is None else sum(weight_count_per_rank[:my_rank])
    local_metadata.shard_sizes = [rows, local_metadata.shard_sizes[1]]
    local_metadata.shard_offsets = [int(offset), local_metadata.shard_offsets[1] if len(local_metadata.shard_offsets) > 1 else 0]


def create_virtual_table_global_metadata(metadata: ShardedTensorMetadata, my_rank: int, param: Union[torch.Tensor, Any],
                                         weight_count_per_rank: Optional[List[int]] = None, use_param_size_as_rows: bool = False) -> None:
    """Row counts of a virtual table differ per rank: rebuild the global metadata from the per-rank counts (reference :204-271)."""
    n = len(metadata.shards_metadata)
    if weight_count_per_rank is None:
        rows_here = int(param.size(0)) if hasattr(param, "size") else int(param)
        weight_count_per_rank = [rows_here if (r == my_rank and use_param_size_as_rows) else (rows_here if use_param_size_as_rows else 1) for r in range(n)]
    off = 0
    for r, sm in enumerate(metadata.shards_metadata):
        cols = sm.shard_sizes[1]
        sm.shard_offsets = [off, sm.shard_offsets[1] if len(sm.shard_offsets) > 1 else 0]
        sm.shard_sizes = [int(weight_count_per_rank[r]), cols]
        off += int(weight_count_per_rank[r])
    metadata.size = torch.Size([off, metadata.size[1]])


def create_virtual_sharded_tensors(embedding_tables: List[ShardedEmbeddingTable], params: List[torch.Tensor], pg: Optional[dist.ProcessGroup] = None,
                                   prefix: str = "", table_name_to_weight_count_per_rank: Optional[Dict[str, List[int]]] = None,
                                   use_param_size_as_rows: bool = False) -> List[ShardedTensor]:
    """ShardedTensors over the materialised rows of key-value tables (reference :274-347)."""
    import copy

    rank = dist.get_rank(pg) if pg is not None and dist.is_initialized() else 0
    out: List[ShardedTensor] = []
    for table, param in zip(embedding_tables, params):
        counts = (table_name_to_weight_count_per_rank or {}).get(table.name)
        local_md = copy.deepcopy(table.local_metadata)
        global_md = copy.deepcopy(table.global_metadata)
        create_virtual_table_local_metadata(local_md, param, rank, weight_count_per_rank=counts)
        if global_md is not None:
            create_virtual_table_global_metadata(global_md, rank, param, counts, use_param_size_as_rows)
        out.append(_make_sharded_tensor([Shard(param, local_md)], global_md, pg))
    return out


def get_state_dict(
    embedding_tables: List[ShardedEmbeddingTable],
    params: Union[nn.ModuleList, List[Union[nn.Module, torch.Tensor]], List[torch.Tensor], List[Tuple[torch.Tensor, ...]]],
    pg: Optional[dist.ProcessGroup] = None,
    destination: Optional[Dict[str, Any]] = None,
    prefix: str = "",
) -> Dict[str, Any]:
    """``{prefix}{table}.weight`` -> Tensor (replicated / unsharded tables) or ShardedTensor (tables with shard metadata).

    Several local shards of one table (column-wise on one rank) are gathered into ONE ShardedTensor (reference :350-447). Quantized kernels pass
    ``(weight, scale_shift)`` tuples; their second halves land under ``{table}.weight_qscale`` / ``weight_qbias`` like the reference."""
    if destination is None:
        destination = OrderedDict()
        destination._metadata = OrderedDict()  # type: ignore[attr-defined]
    key_to_local: Dict[str, List[Shard]] = {}
    key_to_global: Dict[str, Any] = {}
    extra: Dict[str, List[Shard]] = {}

    def weight_of(p: Any) -> torch.Tensor:
        if isinstance(p, nn.Module):
            return p.weight  # type: ignore[return-value]
        return p

    assert len(embedding_tables) == len(params), f"{len(embedding_tables)} tables vs {len(params)} params"
    for table, p in zip(embedding_tables, params):
        key = f"{prefix}{table.name}.weight"
        qparts: Tuple[torch.Tensor, ...] = ()
        if isinstance(p, tuple):
            p, *rest = p
            qparts = tuple(r for r in rest if r is not None)
        w = weight_of(p)
        if table.global_metadata is not None and table.local_metadata is not None:
            key_to_local.setdefault(key, []).append(Shard(w, table.local_metadata))
            key_to_global[key] = table.global_metadata
            for suffix, q in zip(("_qscale", "_qbias") if len(qparts) == 2 else ("_qscaleshift",), qparts):
                extra.setdefault(key + suffix, []).append(Shard(q, table.local_metadata))
        else:
            destination[key] = w
            for suffix, q in zip(("_qscale", "_qbias") if len(qparts) == 2 else ("_qscaleshift",), qparts):
                destination[key + suffix] = q
    for key, shards in key_to_local.items():
        destination[key] = _make_sharded_tensor(shards, key_to_global[key], pg)
    for key, shards in extra.items():
        destination[key] = shards[0].tensor if len(shards) == 1 else [s.tensor for s in shards]
    return destination
