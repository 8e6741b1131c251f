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


# ---- inference: one process drives all devices of a host -----------------------------------------------------------------------------------------------------
def _device_of(device_type: str, rank: int) -> torch.device:
    return torch.device("cpu") if device_type == "cpu" else torch.device(f"{device_type}:{rank}")


class _LocalDeviceRouter:
    """Block partition of the pool rows over the local devices: splits ids by owner, remembers how to put the per-device results back
    into the caller's order."""

    def __init__(self, pool_size: int, world_size: int) -> None:
        self.W = world_size
        self.block = (pool_size + world_size - 1) // world_size

    def split(self, ids: torch.Tensor) -> Tuple[List[torch.Tensor], torch.Tensor]:
        ids = ids.long()
        owner = torch.div(ids, self.block, rounding_mode="floor").clamp(max=self.W - 1)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self.W).tolist()
        sorted_ids = ids[order]
        parts, o = [], 0
        for r in range(self.W):
            parts.append(sorted_ids[o : o + counts[r]] - r * self.block)
            o += counts[r]
        return parts, order

    @staticmethod
    def merge(parts: List[torch.Tensor], order: torch.Tensor, device: torch.device) -> torch.Tensor:
        cat = torch.cat([p.to(device) for p in parts])
        out = torch.empty_like(cat)
        out[order.to(device)] = cat
        return out


class LocalShardPool(nn.Module):
    """One device's rows of an inference tensor pool."""

    def __init__(self, shard: torch.Tensor) -> None:
        super().__init__()
        self.register_buffer("_shard", shard)

    def forward(self, rank_ids: torch.Tensor) -> torch.Tensor:
        return self._shard[rank_ids.to(self._shard.device).long()]

    def update(self, rank_ids: torch.Tensor, values: torch.Tensor) -> None:
        self._shard[rank_ids.to(self._shard.device).long()] = values.to(self._shard.device, self._shard.dtype)


class ShardedInferenceTensorPool(nn.Module):
    """A ``TensorPool`` served by ONE process over the devices of its host (reference tensor_pool.py:324): the rows are block
    partitioned over ``env.world_size`` local devices; ``lookup(ids)`` gathers every device's part and returns the rows on the device of
    ``ids`` in the caller's order. Read-only: pools are filled before they are sharded for serving."""

    def __init__(self, pool: TensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._pool_size, self._dim, self._dtype = pool.pool_size, pool.dim, pool.dtype
        dev_type = (device or pool.device).type
        if dev_type == "cuda" and (not torch.cuda.is_available() or torch.cuda.device_count() < env.world_size):
            dev_type = "cpu"
        self._router = _LocalDeviceRouter(self._pool_size, env.world_size)
        b = self._router.block
        self._local_shard_pools = nn.ModuleList([LocalShardPool(pool.pool[r * b : min((r + 1) * b, self._pool_size)].detach().clone().to(_device_of(dev_type, r)))
                                                 for r in range(env.world_size)])

    @property
    def pool_size(self) -> int:
        return self._pool_size

    @property
    def dim(self) -> int:
        return self._dim

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype

    @torch.no_grad()
    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        parts, order = self._router.split(ids)
        return _LocalDeviceRouter.merge([p(i) for p, i in zip(self._local_shard_pools, parts)], order, ids.device)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return self.lookup(ids)

    def update(self, ids: torch.Tensor, values: torch.Tensor) -> None:
        raise NotImplementedError("Inference does not support update")

    def create_context(self) -> None:
        return None


def _infer_shard(self, module: TensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None):
    """``TensorPoolSharder.shard`` picks the serving form for plans marked ``inference``."""
    return ShardedInferenceTensorPool(module, plan, env, device) if getattr(plan, "inference", False) else ShardedTensorPool(module, plan, env, device)


TensorPoolSharder.shard = _infer_shard  # type: ignore[method-assign]
