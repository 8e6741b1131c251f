"""Row-wise sharded ``TensorPool``: ids travel to the owner of their row block, rows travel back in the caller's order; updates go the
same way (reference ``torchrec/distributed/tensor_pool.py``, ``sharding/rw_tensor_pool_sharding.py``). Routing helper: ``object_pool._Router``."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..modules.object_pool import KeyedJaggedTensorPool, TensorPool  # noqa: F401
from ..sparse.jagged_tensor import KeyedJaggedTensor  # noqa: F401
from .object_pool import ObjectPoolShardingPlan, ObjectPoolShardingType, _Router  # noqa: F401
from .types import ShardingEnv


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


class TensorPoolSharder:
    module_type = TensorPool

    def shard(self, module: TensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None) -> ShardedTensorPool:
        return ShardedTensorPool(module, plan, env, device)
