"""Sharded ITEP embedding collections (reference torchrec/distributed/itep_embeddingbag.py:71-596).

The pruning state (address lookup, row utilisation) of a table lives with the module on every rank that feeds ids: the
remap ``logical row -> physical row`` happens on the *source* rank before the input dist, so any sharding type of the
underlying collection works. Utilisation counters are all-reduced at pruning time so all ranks take the same decision and
keep identical address tables; rows whose owner changed are re-initialised through ``reset_rows``."""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Type

import torch
import torch.distributed as dist
from torch import nn

from ..modules.itep_modules import GenericITEPModule, ITEPEmbeddingBagCollection, ITEPEmbeddingCollection
from ..sparse.jagged_tensor import KeyedJaggedTensor
from .embedding import EmbeddingCollectionSharder
from .embedding_types import BaseEmbeddingSharder
from .embeddingbag import EmbeddingBagCollectionSharder
from .types import ParameterSharding, ShardedModule, ShardingEnv


class _ShardedITEPBase(ShardedModule):
    def __init__(self, module, inner_module, params: Dict[str, ParameterSharding], inner_sharder, env: ShardingEnv, device: torch.device) -> None:
        super().__init__()
        self._env, self._device = env, device
        self._inner = inner_sharder.shard(inner_module, params, env=env, device=device)
        self._itep_module: GenericITEPModule = module._itep_module.to(device)
        self.register_buffer("_iter", torch.tensor(0, dtype=torch.int64), persistent=True)
        self._iter_host = 0

    def create_context(self):
        return self._inner.create_context()

    @torch.no_grad()
    def _maybe_prune(self) -> None:
        it = self._itep_module
        if not (self.training and it.enable_pruning and self._iter_host > 0 and self._iter_host % it.pruning_interval == 0 and it.last_pruned_iter != self._iter_host):
            return
        if self._env.world_size > 1:
            for t in it._tables:
                dist.all_reduce(it._util(t), group=self._env.process_group)
        for table, phys in it.prune().items():
            self._inner.reset_rows(table, phys)
        it.last_pruned_iter = self._iter_host

    def input_dist(self, ctx, features: KeyedJaggedTensor, force_insert: bool = False):
        # interval -1: pruning is driven here (after the cross-rank reduction), not inside the module
        features = self._itep_module(features, -1)
        self._maybe_prune()
        if self.training:
            self._iter_host += 1
            self._iter.fill_(self._iter_host)
        return self._inner.input_dist(ctx, features)

    def compute(self, ctx, dist_input):
        return self._inner.compute(ctx, dist_input)

    def output_dist(self, ctx, output):
        return self._inner.output_dist(ctx, output)

    def compute_and_output_dist(self, ctx, input):
        return self._inner.compute_and_output_dist(ctx, input)

    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        p = prefix + "." if prefix else ""
        yield from self._inner.sharded_parameter_names(p + "_inner")

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        from .types import delegating_named_parameters

        yield from delegating_named_parameters(self, prefix, recurse)

    def state_dict(self, destination=None, prefix: str = "", keep_vars: bool = False):  # type: ignore[override]
        from .types import delegating_state_dict

        return delegating_state_dict(self, destination, prefix, keep_vars)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):  # type: ignore[override]
        from .types import delegating_load_state_dict

        return delegating_load_state_dict(self, state_dict, strict)

    @property
    def fused_optimizer(self):
        return self._inner.fused_optimizer


class ShardedITEPEmbeddingBagCollection(_ShardedITEPBase):
    def __init__(self, module: ITEPEmbeddingBagCollection, params, ebc_sharder, env, device) -> None:
        super().__init__(module, module._embedding_bag_collection, params, ebc_sharder, env, device)

    @property
    def _embedding_bag_collection(self):
        return self._inner


class ShardedITEPEmbeddingCollection(_ShardedITEPBase):
    def __init__(self, module: ITEPEmbeddingCollection, params, ec_sharder, env, device) -> None:
        super().__init__(module, module._embedding_collection, params, ec_sharder, env, device)

    @property
    def _embedding_collection(self):
        return self._inner


class ITEPEmbeddingBagCollectionSharder(BaseEmbeddingSharder[ITEPEmbeddingBagCollection]):
    def __init__(self, ebc_sharder: Optional[EmbeddingBagCollectionSharder] = None, fused_params=None, qcomm_codecs_registry=None) -> None:
        super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._ebc_sharder = ebc_sharder or EmbeddingBagCollectionSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)

    def shard(self, module: ITEPEmbeddingBagCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedITEPEmbeddingBagCollection:
        device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedITEPEmbeddingBagCollection(module, params, self._ebc_sharder, env, device)

    def shardable_parameters(self, module: ITEPEmbeddingBagCollection) -> Dict[str, nn.Parameter]:
        return self._ebc_sharder.shardable_parameters(module._embedding_bag_collection)

    @property
    def module_type(self) -> Type[ITEPEmbeddingBagCollection]:
        return ITEPEmbeddingBagCollection


class ITEPEmbeddingCollectionSharder(BaseEmbeddingSharder[ITEPEmbeddingCollection]):
    def __init__(self, ec_sharder: Optional[EmbeddingCollectionSharder] = None, fused_params=None, qcomm_codecs_registry=None) -> None:
        super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._ec_sharder = ec_sharder or EmbeddingCollectionSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)

    def shard(self, module: ITEPEmbeddingCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedITEPEmbeddingCollection:
        device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedITEPEmbeddingCollection(module, params, self._ec_sharder, env, device)

    def shardable_parameters(self, module: ITEPEmbeddingCollection) -> Dict[str, nn.Parameter]:
        return self._ec_sharder.shardable_parameters(module._embedding_collection)

    @property
    def module_type(self) -> Type[ITEPEmbeddingCollection]:
        return ITEPEmbeddingCollection
