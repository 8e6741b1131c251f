"""Row-wise sharded object pools (reference torchrec/distributed/tensor_pool.py, keyed_jagged_tensor_pool.py,
sharding/rw_tensor_pool_sharding.py, rw_kjt_pool_sharding.py, rw_pool_sharding.py).

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

from ..modules.object_pools import KeyedJaggedTensorPool, TensorPool
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


class ShardedTensorPool(nn.Module):
    def __init__(self, pool: TensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._env, self._plan = env, plan
        self._device = device or pool.device
        self._pool_size, self._dim, self._dtype = pool.pool_size, pool.dim, pool.dtype
        self._replicated = plan.sharding_type == ObjectPoolShardingType.REPLICATED_ROW_WISE or env.world_size == 1
        self._router = _Router(self._pool_size, env, self._device)
        rows = self._pool_size if self._replicated else self._router.local_rows
        lo = 0 if self._replicated else env.rank * self._router.block
        self.register_buffer("_local", pool.pool[lo : lo + rows].detach().clone().to(self._device))

    @property
    def pool_size(self) -> int:
        return self._pool_size

    @property
    def dim(self) -> int:
        return self._dim

    @torch.no_grad()
    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        if self._replicated:
            return self._local[ids.long()]
        local_ids, order, sc, rc = self._router.route(ids)
        return self._router.return_rows(self._local[local_ids], order, sc, rc)

    @torch.no_grad()
    def update(self, ids: torch.Tensor, values: torch.Tensor) -> None:
        assert values.shape[1] == self._dim
        if self._env.world_size == 1:
            self._local[ids.long()] = values.to(self._dtype)
            return
        if self._replicated:
            n = torch.tensor([ids.numel()], device=ids.device)
            ns = [torch.zeros_like(n) for _ in range(self._env.world_size)]
            dist.all_gather(ns, n, group=self._env.process_group)
            mx = int(max(int(x) for x in ns))
            pid = torch.full((mx,), -1, dtype=torch.long, device=ids.device)
            pid[: ids.numel()] = ids.long()
            pv = torch.zeros(mx, self._dim, dtype=self._dtype, device=values.device)
            pv[: ids.numel()] = values.to(self._dtype)
            gi = [torch.empty_like(pid) for _ in ns]
            gv = [torch.empty_like(pv) for _ in ns]
            dist.all_gather(gi, pid, group=self._env.process_group)
            dist.all_gather(gv, pv, group=self._env.process_group)
            for i, v in zip(gi, gv):
                m = i >= 0
                self._local[i[m]] = v[m]
            return
        local_ids, order, sc, rc = self._router.route(ids)
        self._local[local_ids] = self._router.send_rows(values.to(self._dtype), order, sc, rc)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return self.lookup(ids)


class ShardedKeyedJaggedTensorPool(nn.Module):
    def __init__(self, pool: KeyedJaggedTensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._env = env
        self._device = device or pool._device
        self._keys = list(pool._keys)
        self._max = dict(pool.feature_max_lengths)
        self._pool_size = pool.pool_size
        self._weighted = pool._is_weighted
        self._router = _Router(self._pool_size, env, self._device)
        self._replicated = plan.sharding_type == ObjectPoolShardingType.REPLICATED_ROW_WISE or env.world_size == 1
        lo = 0 if self._replicated else env.rank * self._router.block
        rows = self._pool_size if self._replicated else self._router.local_rows
        self.register_buffer("_values", pool._values[lo : lo + rows].detach().clone().to(self._device))
        self.register_buffer("_lengths", pool._lengths[lo : lo + rows].detach().clone().to(self._device))
        self._offsets = list(pool._offsets)

    def _pack(self, ids_local: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._values[ids_local], self._lengths[ids_local]

    def _to_kjt(self, dense: torch.Tensor, lengths: torch.Tensor) -> KeyedJaggedTensor:
        vals = []
        for fi, k in enumerate(self._keys):
            block = dense[:, self._offsets[fi] : self._offsets[fi + 1]]
            mask = torch.arange(block.shape[1], device=block.device).unsqueeze(0) < lengths[:, fi : fi + 1]
            vals.append(block[mask])
        return KeyedJaggedTensor(keys=self._keys, values=torch.cat(vals), lengths=lengths.t().reshape(-1), stride=dense.shape[0])

    @torch.no_grad()
    def lookup(self, ids: torch.Tensor) -> KeyedJaggedTensor:
        if self._replicated:
            return self._to_kjt(*self._pack(ids.long()))
        local_ids, order, sc, rc = self._router.route(ids)
        v, l = self._pack(local_ids)
        return self._to_kjt(self._router.return_rows(v, order, sc, rc), self._router.return_rows(l, order, sc, rc))

    @torch.no_grad()
    def update(self, ids: torch.Tensor, values: KeyedJaggedTensor) -> None:
        jt = values.to_dict()
        n = ids.numel()
        dense = torch.zeros(n, self._offsets[-1], dtype=self._values.dtype, device=self._values.device)
        lengths = torch.zeros(n, len(self._keys), dtype=torch.int64, device=self._values.device)
        for fi, k in enumerate(self._keys):
            f = jt[k]
            dense[:, self._offsets[fi] : self._offsets[fi + 1]] = f.to_padded_dense(self._max[k]).to(dense.dtype)
            lengths[:, fi] = f.lengths().long().clamp(max=self._max[k])
        if self._replicated and self._env.world_size > 1:
            raise NotImplementedError("replicated KJT pools are read-only after sharding (update the unsharded pool and re-shard)")
        if self._env.world_size == 1:
            self._values[ids.long()], self._lengths[ids.long()] = dense, lengths
            return
        local_ids, order, sc, rc = self._router.route(ids)
        self._values[local_ids] = self._router.send_rows(dense, order, sc, rc)
        self._lengths[local_ids] = self._router.send_rows(lengths, order, sc, rc)

    def forward(self, ids: torch.Tensor) -> KeyedJaggedTensor:
        return self.lookup(ids)


class TensorPoolSharder:
    module_type = TensorPool

    def shard(self, module: TensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None) -> ShardedTensorPool:
        return ShardedTensorPool(module, plan, env, device)


class KeyedJaggedTensorPoolSharder:
    module_type = KeyedJaggedTensorPool

    def shard(self, module: KeyedJaggedTensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None) -> ShardedKeyedJaggedTensorPool:
        return ShardedKeyedJaggedTensorPool(module, plan, env, device)
