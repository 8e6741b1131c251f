"""Row-wise (or replicated) sharded ``KeyedJaggedTensorPool`` (reference ``torchrec/distributed/keyed_jagged_tensor_pool.py``,
``sharding/rw_kjt_pool_sharding.py``): per-id jagged feature rows are block-partitioned over the ranks and looked up / updated through
the same id routing as the tensor pool."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..modules.object_pool import KeyedJaggedTensorPool, TensorPool  # noqa: F401
from ..sparse.jagged_tensor import KeyedJaggedTensor  # noqa: F401
from .object_pool import ObjectPoolShardingPlan, ObjectPoolShardingType, _Router  # noqa: F401
from .types import ShardingEnv


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


class KeyedJaggedTensorPoolSharder:
    module_type = KeyedJaggedTensorPool

    def shard(self, module: KeyedJaggedTensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None) -> ShardedKeyedJaggedTensorPool:
        return ShardedKeyedJaggedTensorPool(module, plan, env, device)


class ShardedInferenceKeyedJaggedTensorPool(nn.Module):
    """A ``KeyedJaggedTensorPool`` served by one process over the devices of its host: the per-id jagged rows are block partitioned
    over ``env.world_size`` local devices, ``lookup(ids)`` returns the KJT of the ids in the caller's order on the device of ``ids``.
    Read-only (reference keyed_jagged_tensor_pool.py ``ShardedInferenceKeyedJaggedTensorPool``)."""

    def __init__(self, pool: KeyedJaggedTensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None) -> None:
        super().__init__()
        from .tensor_pool import LocalShardPool, _device_of, _LocalDeviceRouter

        self._keys, self._offsets, self._pool_size = list(pool._keys), list(pool._offsets), pool.pool_size
        dev_type = (device or pool._device).type
        if dev_type == "cuda" and (not torch.cuda.is_available() or torch.cuda.device_count() < env.world_size):
            dev_type = "cpu"
        self._router = _LocalDeviceRouter(self._pool_size, env.world_size)
        b = self._router.block
        rows = lambda t, r: t[r * b : min((r + 1) * b, self._pool_size)].detach().clone().to(_device_of(dev_type, r))  # noqa: E731
        self._value_shards = nn.ModuleList([LocalShardPool(rows(pool._values, r)) for r in range(env.world_size)])
        self._length_shards = nn.ModuleList([LocalShardPool(rows(pool._lengths, r)) for r in range(env.world_size)])

    @property
    def pool_size(self) -> int:
        return self._pool_size

    @torch.no_grad()
    def lookup(self, ids: torch.Tensor) -> KeyedJaggedTensor:
        from .tensor_pool import _LocalDeviceRouter

        parts, order = self._router.split(ids)
        dense = _LocalDeviceRouter.merge([p(i) for p, i in zip(self._value_shards, parts)], order, ids.device)
        lengths = _LocalDeviceRouter.merge([p(i) for p, i in zip(self._length_shards, parts)], order, ids.device)
        vals = []
        for fi in range(len(self._keys)):
            block = dense[:, self._offsets[fi] : self._offsets[fi + 1]]
            vals.append(block[torch.arange(block.shape[1], device=block.device).unsqueeze(0) < lengths[:, fi : fi + 1]])
        return KeyedJaggedTensor(keys=self._keys, values=torch.cat(vals), lengths=lengths.t().reshape(-1), stride=dense.shape[0])

    def forward(self, ids: torch.Tensor) -> KeyedJaggedTensor:
        return self.lookup(ids)

    def update(self, ids: torch.Tensor, values: KeyedJaggedTensor) -> None:
        raise NotImplementedError("Inference does not support update")


def _infer_shard(self, module: KeyedJaggedTensorPool, plan: ObjectPoolShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None):
    return ShardedInferenceKeyedJaggedTensorPool(module, plan, env, device) if getattr(plan, "inference", False) else ShardedKeyedJaggedTensorPool(module, plan, env, device)


KeyedJaggedTensorPoolSharder.shard = _infer_shard  # type: ignore[method-assign]
