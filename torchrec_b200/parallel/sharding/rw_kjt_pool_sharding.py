"""Row-wise sharding of a KeyedJaggedTensorPool: per id a small jagged record (``F`` features, bounded lengths).

Reference: ``torchrec/distributed/sharding/rw_kjt_pool_sharding.py`` - ``RwKeyedJaggedTensorPoolLookupValuesDist`` :38, ``RwKeyedJaggedTensorPoolUpdateValuesDist`` :105,
``KeyedJaggedTensorPoolRwSharding`` :249, ``InferRwKeyedJaggedTensorPoolOutputDist`` :346, ``InferRwKeyedJaggedTensorPoolSharding`` :409,
``KeyedJaggedTensorPoolRwReplicatedSharding`` :429. Jagged records move as (lengths ``[n, F]``, padded values ``[n, sum(max_len)]``): fixed-width rows make
the exchange two plain all-to-alls, the pool's lookup already stores them this way (``modules/object_pool_lookups.py``).
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ...sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from ..tensor_sharding import InferObjectPoolSharding, ObjectPoolRwShardingContext, ObjectPoolSharding
from ..types import Awaitable, NoWait, ShardingEnv
from .rw_pool_sharding import InferRwObjectPoolInputDist, RwObjectPoolIDsDist


def _pad(jt_values: torch.Tensor, lengths_nf: torch.Tensor, max_lens: List[int], feature_major_lengths: torch.Tensor) -> torch.Tensor:
    """(feature, id)-ordered jagged values -> ``[n, sum(max_len)]`` padded rows."""
    n, F = lengths_nf.shape
    out = torch.zeros(n, sum(max_lens), dtype=jt_values.dtype, device=jt_values.device)
    src = torch.cumsum(feature_major_lengths.long(), 0) - feature_major_lengths.long()
    col = 0
    for f in range(F):
        L = lengths_nf[:, f].long().clamp(max=max_lens[f])
        pos = torch.arange(max_lens[f], device=out.device).unsqueeze(0)
        mask = pos < L.unsqueeze(1)
        idx = (src[f * n : (f + 1) * n].unsqueeze(1) + pos)[mask]
        out[:, col : col + max_lens[f]][mask] = jt_values[idx]
        col += max_lens[f]
    return out


def _unpad(rows: torch.Tensor, lengths_nf: torch.Tensor, max_lens: List[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    vals, lens, col = [], [], 0
    for f in range(lengths_nf.shape[1]):
        L = lengths_nf[:, f].long()
        mask = torch.arange(max_lens[f], device=rows.device).unsqueeze(0) < L.unsqueeze(1)
        vals.append(rows[:, col : col + max_lens[f]][mask])
        lens.append(lengths_nf[:, f])
        col += max_lens[f]
    return (torch.cat(vals) if vals else rows.new_empty(0)), (torch.cat(lens) if lens else lengths_nf.new_empty(0))


class RwKeyedJaggedTensorPoolLookupValuesDist(nn.Module):
    """Looked-up records back to the requesters: ``JaggedTensor`` ((feature, id)-ordered lengths) -> KJT in request order."""

    def __init__(self, num_features: int, env: ShardingEnv, feature_max_lengths: Optional[Dict[str, int]] = None) -> None:
        super().__init__()
        self._pg = env.process_group
        self._num_features = num_features
        self._max_lens = list(feature_max_lengths.values()) if feature_max_lengths else None
        self._keys = list(feature_max_lengths.keys()) if feature_max_lengths else [f"f{i}" for i in range(num_features)]

    def forward(self, ctx: ObjectPoolRwShardingContext, jagged_tensor: JaggedTensor) -> Awaitable[KeyedJaggedTensor]:
        F = self._num_features
        n = sum(ctx.recv_counts)
        lens_nf = jagged_tensor.lengths().view(F, n).t().contiguous()
        max_lens = self._max_lens or [int(lens_nf[:, f].max()) if n else 0 for f in range(F)]
        rows = _pad(jagged_tensor.values(), lens_nf, max_lens, jagged_tensor.lengths())
        m = sum(ctx.send_counts)
        back_l = torch.empty(m, F, dtype=lens_nf.dtype, device=lens_nf.device)
        back_v = torch.empty(m, rows.shape[1], dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(back_l, lens_nf, output_split_sizes=ctx.send_counts, input_split_sizes=ctx.recv_counts, group=self._pg)
        dist.all_to_all_single(back_v, rows, output_split_sizes=ctx.send_counts, input_split_sizes=ctx.recv_counts, group=self._pg)
        perm = ctx.unbucketize_permute
        back_l, back_v = back_l[perm], back_v[perm]
        vals, lens = _unpad(back_v, back_l, max_lens)
        return NoWait(KeyedJaggedTensor(keys=self._keys, values=vals, lengths=lens, stride=m))


class RwKeyedJaggedTensorPoolUpdateValuesDist(nn.Module):
    """New records to the owners: KJT (one sample per id) -> JaggedTensor in the owner's received-id order."""

    def __init__(self, num_features: int, env: ShardingEnv, device: torch.device, num_replicas: int = 1, feature_max_lengths: Optional[Dict[str, int]] = None) -> None:
        super().__init__()
        self._pg = env.process_group
        self._num_features = num_features
        self._max_lens = list(feature_max_lengths.values()) if feature_max_lengths else None

    def forward(self, ctx: ObjectPoolRwShardingContext, values: KeyedJaggedTensor) -> Awaitable[JaggedTensor]:
        F = self._num_features
        n = values.stride()
        lens_nf = values.lengths().view(F, n).t().contiguous()
        max_lens = self._max_lens or [int(lens_nf[:, f].max()) if n else 0 for f in range(F)]
        rows = _pad(values.values(), lens_nf, max_lens, values.lengths())
        lens_nf, rows = lens_nf[ctx.order].contiguous(), rows[ctx.order].contiguous()
        m = sum(ctx.recv_counts)
        out_l = torch.empty(m, F, dtype=lens_nf.dtype, device=lens_nf.device)
        out_v = torch.empty(m, rows.shape[1], dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out_l, lens_nf, output_split_sizes=ctx.recv_counts, input_split_sizes=ctx.send_counts, group=self._pg)
        dist.all_to_all_single(out_v, rows, output_split_sizes=ctx.recv_counts, input_split_sizes=ctx.send_counts, group=self._pg)
        vals, lens = _unpad(out_v, out_l, max_lens)
        return NoWait(JaggedTensor(values=vals, lengths=lens))


class KeyedJaggedTensorPoolRwSharding(ObjectPoolSharding):
    def __init__(self, env: ShardingEnv, device: torch.device, pool_size: int, num_features: int, feature_max_lengths: Optional[Dict[str, int]] = None) -> None:
        self._env, self._pg, self._device = env, env.process_group, device
        self._world_size, self._rank = env.world_size, env.rank
        self._pool_size, self._num_features = pool_size, num_features
        self._feature_max_lengths = feature_max_lengths
        self._block_size = (pool_size + self._world_size - 1) // self._world_size
        self.local_pool_size = max(0, min(self._block_size, pool_size - self._rank * self._block_size))
        self._block_size_t = torch.tensor([self._block_size], device=device, dtype=torch.long)

    def create_update_ids_dist(self) -> RwObjectPoolIDsDist:
        return RwObjectPoolIDsDist(self._pg, is_update=True)

    def create_update_values_dist(self) -> RwKeyedJaggedTensorPoolUpdateValuesDist:
        return RwKeyedJaggedTensorPoolUpdateValuesDist(self._num_features, self._env, self._device, feature_max_lengths=self._feature_max_lengths)

    def create_lookup_ids_dist(self) -> RwObjectPoolIDsDist:
        return RwObjectPoolIDsDist(self._pg, is_update=False)

    def create_lookup_values_dist(self) -> RwKeyedJaggedTensorPoolLookupValuesDist:
        return RwKeyedJaggedTensorPoolLookupValuesDist(self._num_features, self._env, self._feature_max_lengths)

    def get_sharded_states_to_register(self, lookup: nn.Module) -> Iterator[Tuple[str, torch.Tensor]]:
        yield from lookup.states_to_register()

    def create_context(self) -> ObjectPoolRwShardingContext:
        return ObjectPoolRwShardingContext(block_size=self._block_size_t)


class KeyedJaggedTensorPoolRwReplicatedSharding(KeyedJaggedTensorPoolRwSharding):
    """Pool replicated per HOST, row-wise inside a host: lookups never leave the NVLink domain; updates go to the owner in EVERY host.

    ``env`` is the global env; ids are bucketized over ``local_world_size`` ranks and offset to this rank's host for lookups."""

    def __init__(self, env: ShardingEnv, device: torch.device, pool_size: int, num_features: int, feature_max_lengths: Optional[Dict[str, int]] = None,
                 local_world_size: Optional[int] = None) -> None:
        from ..comm import get_local_size

        super().__init__(env, device, pool_size, num_features, feature_max_lengths)
        self._local_world_size = local_world_size or get_local_size(env.world_size)
        self._num_replicas = self._world_size // self._local_world_size
        self._block_size = (pool_size + self._local_world_size - 1) // self._local_world_size
        local_rank = self._rank % self._local_world_size
        self.local_pool_size = max(0, min(self._block_size, pool_size - local_rank * self._block_size))
        self._block_size_t = torch.tensor([self._block_size], device=device, dtype=torch.long)

    def create_lookup_ids_dist(self) -> RwObjectPoolIDsDist:
        host0 = (self._rank // self._local_world_size) * self._local_world_size
        return RwObjectPoolIDsDist(self._pg, is_update=False, bucketize_world_size=self._local_world_size, bucketize_rank_offset=host0)


class InferRwKeyedJaggedTensorPoolOutputDist(nn.Module):
    """Merge the per-device JaggedTensors into one KJT in request order."""

    def __init__(self, env: ShardingEnv, device: torch.device) -> None:
        super().__init__()
        self._world_size = env.world_size
        self._device = device

    def forward(self, jagged_tensors: List[JaggedTensor], keys: List[str], unbucketize_permute: torch.Tensor) -> KeyedJaggedTensor:
        F = len(keys)
        lens_nf = torch.cat([jt.lengths().to(self._device).view(F, -1).t() for jt in jagged_tensors], dim=0)
        max_lens = [int(lens_nf[:, f].max()) if lens_nf.numel() else 0 for f in range(F)]
        rows = torch.cat([_pad(jt.values().to(self._device), jt.lengths().to(self._device).view(F, -1).t().contiguous(), max_lens, jt.lengths().to(self._device)) for jt in jagged_tensors], dim=0)
        perm = unbucketize_permute.to(self._device)
        vals, lens = _unpad(rows[perm], lens_nf[perm], max_lens)
        return KeyedJaggedTensor(keys=keys, values=vals, lengths=lens, stride=lens_nf.shape[0])


class InferRwKeyedJaggedTensorPoolSharding(InferObjectPoolSharding):
    def create_lookup_ids_dist(self) -> InferRwObjectPoolInputDist:
        return InferRwObjectPoolInputDist(self._env, self._device, self._block_size_t)

    def create_lookup_values_dist(self) -> InferRwKeyedJaggedTensorPoolOutputDist:
        return InferRwKeyedJaggedTensorPoolOutputDist(self._env, self._device)
