"""Sharded managed-collision collection (reference torchrec/distributed/mc_modules.py:208-1700).

The id space of every ZCH table is partitioned over the ranks: rank r owns output slots ``[seg[r], seg[r+1])`` and the
raw ids that hash to it. ``forward`` is a remap *service*: raw ids travel to their owner (all-to-all), the owner's local
ManagedCollisionModule (rebuilt with ``rebuild_with_output_id_range``) turns them into global slot ids, and the slot ids
travel back into the original KJT positions. Unlike the reference — which only supports row-wise tables laid out exactly
like the MC partition — the remapped KJT here is an ordinary KJT in the global slot space, so the embedding collection
behind it can use ANY sharding type of the lookup-unit engine. Evicted slots are all-gathered so that whichever rank
holds the rows can re-initialise them."""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Tuple, Type

import torch
import torch.distributed as dist
from torch import nn

from ..modules.mc_modules import ManagedCollisionCollection, ManagedCollisionModule
from ..ops import jagged as J
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .types import ModuleSharder, ParameterSharding, ShardingEnv, ShardingType


def _owner_of(ids: torch.Tensor, W: int) -> torch.Tensor:
    """Owner rank of a raw id: multiplicative hash so that dense small ids spread evenly."""
    x = ids.to(torch.int64)
    x = (x ^ (x >> 31)) * -7046029254386353131  # 0x9e3779b97f4a7c15
    x = x ^ (x >> 29)
    return torch.remainder(x, W)


class ShardedManagedCollisionCollection(nn.Module):
    def __init__(self, module: ManagedCollisionCollection, table_name_to_parameter_sharding: Dict[str, ParameterSharding], env: ShardingEnv,
                 device: torch.device, embedding_shardings: Optional[List[Any]] = None, use_index_dedup: bool = False) -> None:
        super().__init__()
        self._env = env
        self._pg = env.process_group
        self._W = env.world_size
        self._rank = env.rank
        self._device = device
        self._embedding_configs = module.embedding_configs()
        self._table_to_features = {c.name: list(c.feature_names) for c in self._embedding_configs}
        self._features_order: List[str] = [f for c in self._embedding_configs for f in c.feature_names]
        self._feature_table = [c.name for c in self._embedding_configs for _ in c.feature_names]
        mods: Dict[str, ManagedCollisionModule] = {}
        self._segments: Dict[str, List[int]] = {}
        for name, mc in module._managed_collision_modules.items():
            zch = mc.output_size()
            W = self._W
            buckets = mc.buckets() if hasattr(mc, "buckets") else W
            if buckets and buckets % W == 0 and zch % buckets == 0 and buckets != W:
                per = zch // buckets
                segs_b = [per * i for i in range(buckets + 1)]
                lo, hi = segs_b[self._rank * (buckets // W)], segs_b[(self._rank + 1) * (buckets // W)]
                mods[name] = mc.rebuild_with_output_id_range((lo, hi), segs_b, device)
                self._segments[name] = [segs_b[r * (buckets // W)] for r in range(W + 1)]
            else:
                block = (zch + W - 1) // W
                segs = [min(block * r, zch) for r in range(W + 1)]
                mods[name] = mc.rebuild_with_output_id_range((segs[self._rank], segs[self._rank + 1]), segs, device)
                self._segments[name] = segs
        self._managed_collision_modules = nn.ModuleDict(mods)
        self._pending_evictions: Dict[str, torch.Tensor] = {}

    def embedding_configs(self):
        return self._embedding_configs

    # ---- remap round trip ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        W, F = self._W, len(self._features_order)
        if features.keys() != self._features_order:
            features = features.permute([features.keys().index(f) for f in self._features_order])
        B = features.stride()
        values, lengths = features.values(), features.lengths().to(torch.int64)
        n = values.numel()
        dev = values.device
        if W == 1:
            return self._remap_local(features)
        bag = torch.repeat_interleave(torch.arange(F * B, device=dev), lengths, output_size=n)
        owner = _owner_of(values, W)
        new_bag = owner * (F * B) + bag
        order = torch.argsort(new_bag, stable=True)
        send_lengths = torch.bincount(new_bag, minlength=W * F * B)
        send_values = values[order].contiguous()
        send_counts = send_lengths.view(W, -1).sum(1)
        # exchange batch sizes (variable batch per rank), lengths, then values
        Bs = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(W)]
        dist.all_gather(Bs, torch.tensor([B], dtype=torch.int64, device=dev), group=self._pg)
        B_per_rank = [int(b) for b in Bs]
        recv_lengths = torch.empty(F * sum(B_per_rank), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv_lengths, send_lengths, output_split_sizes=[F * b for b in B_per_rank], input_split_sizes=[F * B] * W, group=self._pg)
        seg_ends = torch.tensor([F * b for b in B_per_rank], device=dev).cumsum(0)
        csum = torch.cat([recv_lengths.new_zeros(1), recv_lengths.cumsum(0)])
        recv_counts_t = csum[seg_ends] - csum[torch.cat([seg_ends.new_zeros(1), seg_ends[:-1]])]
        cnt = torch.stack([send_counts, recv_counts_t]).tolist()
        send_counts_l, recv_counts_l = [int(x) for x in cnt[0]], [int(x) for x in cnt[1]]
        recv_values = torch.empty(sum(recv_counts_l), dtype=values.dtype, device=dev)
        dist.all_to_all_single(recv_values, send_values, output_split_sizes=recv_counts_l, input_split_sizes=send_counts_l, group=self._pg)
        # rank-major [r][f][b]  ->  per-feature jagged tensors over all sources
        remapped = torch.empty_like(recv_values)
        seg_len = []  # per (r, f) value counts
        pos = 0
        bounds: List[Tuple[int, int, int]] = []  # (r, f, bag_start)
        for r in range(W):
            for f in range(F):
                bounds.append((r, f, pos))
                pos += B_per_rank[r]
        bag_starts = torch.tensor([b[2] for b in bounds] + [pos], device=dev)
        seg_off = csum[bag_starts]  # value offset of every (r, f) segment
        seg_off_l = seg_off.tolist()
        per_table_feats: Dict[str, Dict[str, Tuple[torch.Tensor, torch.Tensor, List[Tuple[int, int]]]]] = {}
        for f, fname in enumerate(self._features_order):
            spans = [(int(seg_off_l[r * F + f]), int(seg_off_l[r * F + f + 1])) for r in range(W)]
            vals = torch.cat([recv_values[a:b] for a, b in spans])
            lens = torch.cat([recv_lengths[bounds[r * F + f][2] : bounds[r * F + f][2] + B_per_rank[r]] for r in range(W)])
            per_table_feats.setdefault(self._feature_table[f], {})[fname] = (vals, lens, spans)
        for table, feats in per_table_feats.items():
            mc = self._managed_collision_modules[table]
            out = mc({k: JaggedTensor(values=v, lengths=l) for k, (v, l, _) in feats.items()})
            for k, (_, _, spans) in feats.items():
                ov, c = out[k].values(), 0
                for a, b in spans:
                    remapped[a:b] = ov[c : c + (b - a)].to(remapped.dtype)
                    c += b - a
        back = torch.empty(n, dtype=values.dtype, device=dev)
        dist.all_to_all_single(back, remapped, output_split_sizes=send_counts_l, input_split_sizes=recv_counts_l, group=self._pg)
        new_values = torch.empty_like(back)
        new_values[order] = back
        return KeyedJaggedTensor(keys=self._features_order, values=new_values, lengths=features.lengths(), weights=features.weights_or_none(), stride=B)

    def _remap_local(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        jt = features.to_dict()
        out: Dict[str, JaggedTensor] = {}
        for table, mc in self._managed_collision_modules.items():
            out.update(mc({f: jt[f] for f in self._table_to_features[table]}))
        keys = self._features_order
        return KeyedJaggedTensor(keys=keys, values=torch.cat([out[k].values() for k in keys]), lengths=features.lengths(), weights=features.weights_or_none(),
                                 stride=features.stride())

    # ---- evictions ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def evict(self) -> Dict[str, Optional[torch.Tensor]]:
        """Global slot ids evicted on ANY rank since the last call (every rank gets the full list)."""
        res: Dict[str, Optional[torch.Tensor]] = {}
        for table, mc in self._managed_collision_modules.items():
            ev = mc.evict()
            ev = ev.to(self._device).to(torch.int64) if ev is not None else torch.zeros(0, dtype=torch.int64, device=self._device)
            if self._W > 1:
                sizes = [torch.zeros(1, dtype=torch.int64, device=self._device) for _ in range(self._W)]
                dist.all_gather(sizes, torch.tensor([ev.numel()], dtype=torch.int64, device=self._device), group=self._pg)
                sizes_l = [int(s) for s in sizes]
                if sum(sizes_l) == 0:
                    res[table] = None
                    continue
                mx = max(sizes_l)
                pad = torch.full((mx,), -1, dtype=torch.int64, device=self._device)
                pad[: ev.numel()] = ev
                outs = [torch.empty_like(pad) for _ in range(self._W)]
                dist.all_gather(outs, pad, group=self._pg)
                ev = torch.cat([o[:s] for o, s in zip(outs, sizes_l)])
            res[table] = ev if ev.numel() else None
        return res

    def open_slots(self) -> Dict[str, torch.Tensor]:
        res = {}
        for t, mc in self._managed_collision_modules.items():
            s = mc.open_slots().to(self._device)
            if self._W > 1:
                dist.all_reduce(s, group=self._pg)
            res[t] = s
        return res

    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        return iter(())


class ManagedCollisionCollectionSharder(ModuleSharder[ManagedCollisionCollection]):
    def shard(self, module: ManagedCollisionCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, embedding_shardings: Optional[List[Any]] = None,
              device: Optional[torch.device] = None, use_index_dedup: bool = False, module_fqn: Optional[str] = None) -> ShardedManagedCollisionCollection:
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedManagedCollisionCollection(module, params, env, device, embedding_shardings, use_index_dedup)

    def shardable_parameters(self, module: ManagedCollisionCollection) -> Dict[str, nn.Parameter]:
        return {}

    @property
    def module_type(self) -> Type[ManagedCollisionCollection]:
        return ManagedCollisionCollection

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [ShardingType.ROW_WISE.value]
