"""Shared base of the sharded managed-collision embedding collections
(reference torchrec/distributed/mc_embedding_modules.py:62-416)."""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Tuple, Type, Union

import torch
from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor
from .mc_modules import ManagedCollisionCollectionSharder, ShardedManagedCollisionCollection
from .types import Awaitable, LazyAwaitable, Multistreamable, NoWait, ParameterSharding, ShardedModule, ShardingEnv


class ManagedCollisionCollectionContext(Multistreamable):
    def __init__(self, inner: Any = None) -> None:
        self.inner = inner
        self.remapped_kjt: Optional[KeyedJaggedTensor] = None

    def record_stream(self, stream: torch.Stream) -> None:
        if self.inner is not None:
            self.inner.record_stream(stream)
        if self.remapped_kjt is not None:
            self.remapped_kjt.record_stream(stream)


class _WithRemapped(LazyAwaitable):
    def __init__(self, inner: Awaitable, remapped: Optional[KeyedJaggedTensor]) -> None:
        super().__init__()
        self._inner, self._remapped = inner, remapped

    def _wait_impl(self):
        return self._inner.wait(), self._remapped


class BaseShardedManagedCollisionEmbeddingCollection(ShardedModule):
    """input_dist = remap round trip + the embedding module's input dist; compute / output_dist delegate."""

    def __init__(self, module, table_name_to_parameter_sharding: Dict[str, ParameterSharding], e_sharder, mc_sharder: ManagedCollisionCollectionSharder,
                 env: ShardingEnv, device: torch.device, module_fqn: Optional[str] = None) -> None:
        super().__init__()
        self._device = device
        self._env = env
        self._embedding_module = e_sharder.shard(module._embedding_module, table_name_to_parameter_sharding, env=env, device=device)
        self._managed_collision_collection: ShardedManagedCollisionCollection = mc_sharder.shard(
            module._managed_collision_collection, table_name_to_parameter_sharding, env=env, device=device)
        self._return_remapped_features: bool = module._return_remapped_features
        self._table_cfg = {c.name: c for c in module._managed_collision_collection.embedding_configs()}

    def create_context(self) -> ManagedCollisionCollectionContext:
        return ManagedCollisionCollectionContext(self._embedding_module.create_context())

    @torch.no_grad()
    def _evict(self) -> None:
        for table, ids in self._managed_collision_collection.evict().items():
            if ids is not None and ids.numel():
                self._embedding_module.reset_rows(table, ids)

    def input_dist(self, ctx: ManagedCollisionCollectionContext, features: KeyedJaggedTensor):
        remapped = self._managed_collision_collection(features)
        if self.training:
            self._evict()
        if self._return_remapped_features:
            ctx.remapped_kjt = remapped
        return self._embedding_module.input_dist(ctx.inner, remapped)

    def compute(self, ctx: ManagedCollisionCollectionContext, dist_input):
        return self._embedding_module.compute(ctx.inner, dist_input)

    def output_dist(self, ctx: ManagedCollisionCollectionContext, output):
        return _WithRemapped(self._embedding_module.output_dist(ctx.inner, output), ctx.remapped_kjt)

    def compute_and_output_dist(self, ctx: ManagedCollisionCollectionContext, input):
        return _WithRemapped(self._embedding_module.compute_and_output_dist(ctx.inner, input), ctx.remapped_kjt)

    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        p = prefix + "." if prefix else ""
        yield from self._embedding_module.sharded_parameter_names(p + "_embedding_module")

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        from .types import delegating_named_parameters

        yield from delegating_named_parameters(self, prefix, recurse)

    @property
    def fused_optimizer(self):
        return self._embedding_module.fused_optimizer
