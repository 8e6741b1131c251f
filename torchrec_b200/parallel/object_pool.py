"""Common parts of the row-wise sharded object pools: sharding type / plan and the id router (reference torchrec/distributed/
sharding/rw_pool_sharding.py); the sharded pools themselves are in ``tensor_pool.py`` and ``keyed_jagged_tensor_pool.py``.

Pool rows are block-partitioned over the ranks (``ObjectPoolShardingType.ROW_WISE``) or fully replicated
(``REPLICATED_ROW_WISE``: every rank holds the whole pool, updates are all-gathered). ``lookup(ids)``: ids travel to their
owner (all-to-all), the owner gathers the rows, the rows travel back into the caller's order. ``update(ids, values)``: ids
and rows travel to the owner, which writes them."""
from __future__ import annotations

from enum import Enum, unique
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..modules.object_pool import KeyedJaggedTensorPool, TensorPool
from ..sparse.jagged_tensor import KeyedJaggedTensor
from .types import ShardingEnv


@unique
class ObjectPoolShardingType(Enum):
    ROW_WISE = "row_wise"
    REPLICATED_ROW_WISE = "replicated_row_wise"


class ObjectPoolShardingPlan:
    def __init__(self, sharding_type: ObjectPoolShardingType = ObjectPoolShardingType.ROW_WISE, inference: bool = False) -> None:
        self.sharding_type = sharding_type
        self.inference = inference


class _Router:
    """Sends ids (and optional per-id payload rows) to the block owner and brings per-id results back in order."""

    def __init__(self, pool_size: int, env: ShardingEnv, device: torch.device) -> None:
        self.W, self.rank, self.pg, self.device = env.world_size, env.rank, env.process_group, device
        self.block = (pool_size + self.W - 1) // self.W
        self.local_rows = max(0, min(self.block, pool_size - self.rank * self.block))

    def route(self, ids: torch.Tensor):
        ids = ids.long()
        owner = torch.div(ids, self.block, rounding_mode="floor").clamp(max=self.W - 1)
        order = torch.argsort(owner, stable=True)
        send_counts = torch.bincount(owner, minlength=self.W)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.pg)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        sent = ids[order].contiguous()
        recv = torch.empty(sum(rc), dtype=ids.dtype, device=ids.device)
        dist.all_to_all_single(recv, sent, output_split_sizes=rc, input_split_sizes=sc, group=self.pg)
        return recv - self.rank * self.block, order, sc, rc

    def send_rows(self, rows: torch.Tensor, order: torch.Tensor, sc: List[int], rc: List[int]) -> torch.Tensor:
        out = torch.empty(sum(rc), *rows.shape[1:], dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out, rows[order].contiguous(), output_split_sizes=rc, input_split_sizes=sc, group=self.pg)
        return out

    def return_rows(self, rows: torch.Tensor, order: torch.Tensor, sc: List[int], rc: List[int]) -> torch.Tensor:
        back = torch.empty(sum(sc), *rows.shape[1:], dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(back, rows.contiguous(), output_split_sizes=sc, input_split_sizes=rc, group=self.pg)
        out = torch.empty_like(back)
        out[order] = back
        return out


# ---- the sharded pools live in ``tensor_pool.py`` / ``keyed_jagged_tensor_pool.py`` (their reference import paths) ----
def __getattr__(name: str):
    if name in ("ShardedTensorPool", "TensorPoolSharder"):
        from . import tensor_pool as _m

        return getattr(_m, name)
    if name in ("ShardedKeyedJaggedTensorPool", "KeyedJaggedTensorPoolSharder"):
        from . import keyed_jagged_tensor_pool as _m

        return getattr(_m, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
