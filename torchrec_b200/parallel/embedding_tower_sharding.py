"""Sharded embedding towers (reference torchrec/distributed/embedding_tower_sharding.py:96-1040).

A tower = an embedding module + the interaction that consumes only that module's output. The reference pins a tower's tables to
one host, shards them inside the host and runs the interaction on that host, so that the cross-host all-to-all carries the (small)
interaction output instead of the pooled embeddings. On an NVSwitch node every rank is "intra-node": the embedding module is
sharded by the ordinary lookup-unit engine (any sharding type, fused NVLink lookup + dist), and the interaction runs data-parallel
on every rank over its local batch slice (its parameters are replicated and reduced by DMP's DDP wrap like any dense parameter).
The numerical contract is the reference's: ``tower(features) == interaction(embedding(features))`` on each rank's batch, with the
unsharded module's state-dict keys (``embedding.*`` sharded, ``interaction.*`` plain)."""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Tuple, Type

import torch
from torch import nn

from ..modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from ..modules.embedding_tower import EmbeddingTower, EmbeddingTowerCollection, tower_input_params
from ..sparse.jagged_tensor import KeyedJaggedTensor
from .embedding import EmbeddingCollectionSharder
from .embedding_types import BaseEmbeddingSharder
from .embeddingbag import EmbeddingBagCollectionSharder
from .types import LazyAwaitable, ParameterSharding, ShardedModule, ShardingEnv, ShardingType


class TowerLazyAwaitable(LazyAwaitable[torch.Tensor]):
    """Applies the tower's interaction when the embedding output is first needed."""

    def __init__(self, awaitable: Any, interaction: nn.Module) -> None:
        super().__init__()
        self._awaitable = awaitable
        self._interaction = interaction

    def _wait_impl(self) -> torch.Tensor:
        emb = self._awaitable.wait() if hasattr(self._awaitable, "wait") else self._awaitable
        return self._interaction(emb)


def _inner_sharder_for(embedding: nn.Module, ebc_sharder, ec_sharder):
    if isinstance(embedding, EmbeddingBagCollection):
        return ebc_sharder
    if isinstance(embedding, EmbeddingCollection):
        return ec_sharder
    raise TypeError(f"EmbeddingTower: unsupported embedding module {type(embedding).__name__}")


class ShardedEmbeddingTower(ShardedModule):
    def __init__(self, module: EmbeddingTower, params: Dict[str, ParameterSharding], inner_sharder, env: ShardingEnv, device: torch.device) -> None:
        super().__init__()
        self._env, self._device = env, device
        self.embedding = inner_sharder.shard(module.embedding, params, env=env, device=device)
        if any(p.is_meta for p in module.interaction.parameters()):
            # DMP materialises meta parameters of plain modules only; the interaction lives inside this sharded module
            self.interaction = module.interaction.to_empty(device=device)
            for m in self.interaction.modules():
                if hasattr(m, "reset_parameters"):
                    m.reset_parameters()
        else:
            self.interaction = module.interaction.to(device)

    def create_context(self):
        return self.embedding.create_context()

    def input_dist(self, ctx, *input, **kwargs):
        return self.embedding.input_dist(ctx, *input, **kwargs)

    def compute(self, ctx, dist_input):
        return self.embedding.compute(ctx, dist_input)

    def output_dist(self, ctx, output) -> TowerLazyAwaitable:
        return TowerLazyAwaitable(self.embedding.output_dist(ctx, output), self.interaction)

    def compute_and_output_dist(self, ctx, input) -> TowerLazyAwaitable:
        return TowerLazyAwaitable(self.embedding.compute_and_output_dist(ctx, input), self.interaction)

    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        p = prefix + "." if prefix else ""
        yield from self.embedding.sharded_parameter_names(p + "embedding")

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
        return self.embedding.fused_optimizer


class ShardedEmbeddingTowerCollection(nn.Module):
    """Every tower sharded on its own; ``forward(features, weighted_features)`` concatenates the tower outputs like the unsharded
    collection. (A plain module, not a ShardedModule: each tower keeps its own 3-phase pipeline, so train pipelines overlap the
    towers' input dists individually.)"""

    def __init__(self, module: EmbeddingTowerCollection, params: Dict[str, ParameterSharding], ebc_sharder, ec_sharder, env: ShardingEnv, device: torch.device) -> None:
        super().__init__()
        towers = []
        self._input_params: List[Tuple[bool, bool]] = []
        for tower in module.towers:
            inner = _inner_sharder_for(tower.embedding, ebc_sharder, ec_sharder)
            names = set(inner.shardable_parameters(tower.embedding))
            towers.append(ShardedEmbeddingTower(tower, {k: v for k, v in params.items() if k in names}, inner, env, device))
            self._input_params.append(tower_input_params(tower.embedding))
        self.towers = nn.ModuleList(towers)

    def forward(self, features: Optional[KeyedJaggedTensor] = None, weighted_features: Optional[KeyedJaggedTensor] = None) -> torch.Tensor:
        pending = []
        for tower, (has_kjt, has_wkjt) in zip(self.towers, self._input_params):
            if has_kjt and has_wkjt:
                pending.append(tower(features, weighted_features))
            elif has_wkjt:
                pending.append(tower(weighted_features))
            else:
                pending.append(tower(features))
        # all lookups / dists are in flight before the first interaction runs
        return torch.cat([p.wait() if hasattr(p, "wait") else p for p in pending], dim=1)

    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        p = prefix + "." if prefix else ""
        for i, t in enumerate(self.towers):
            yield from t.sharded_parameter_names(f"{p}towers.{i}")

    @property
    def fused_optimizer(self):
        from ..optim.keyed import CombinedOptimizer

        return CombinedOptimizer([(f"towers.{i}", t.fused_optimizer) for i, t in enumerate(self.towers)])


class EmbeddingTowerSharder(BaseEmbeddingSharder[EmbeddingTower]):
    def __init__(self, fused_params=None, qcomm_codecs_registry=None) -> None:
        super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._ebc_sharder = EmbeddingBagCollectionSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._ec_sharder = EmbeddingCollectionSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)

    def embedding_sharder(self, module: EmbeddingTower):
        return _inner_sharder_for(module.embedding, self._ebc_sharder, self._ec_sharder)

    def shard(self, module: EmbeddingTower, params: Dict[str, ParameterSharding], env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedEmbeddingTower:
        device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedEmbeddingTower(module, params, self.embedding_sharder(module), env, device)

    def sharding_types(self, compute_device_type: str) -> List[str]:
        # the reference restricts towers to host-local types; on one NVSwitch domain every type is host-local
        return [ShardingType.TABLE_WISE.value, ShardingType.ROW_WISE.value, ShardingType.COLUMN_WISE.value, ShardingType.TABLE_ROW_WISE.value,
                ShardingType.TABLE_COLUMN_WISE.value]

    def shardable_parameters(self, module: EmbeddingTower) -> Dict[str, nn.Parameter]:
        return self.embedding_sharder(module).shardable_parameters(module.embedding)

    def embedding_feature_names(self, module: EmbeddingTower) -> Tuple[List[str], List[str]]:
        has_kjt, has_wkjt = tower_input_params(module.embedding)
        cfgs = module.embedding.embedding_bag_configs() if isinstance(module.embedding, EmbeddingBagCollection) else module.embedding.embedding_configs()
        names = [f for c in cfgs for f in c.feature_names]
        return (names if has_kjt else []), (names if has_wkjt else [])

    @property
    def module_type(self) -> Type[EmbeddingTower]:
        return EmbeddingTower


class EmbeddingTowerCollectionSharder(BaseEmbeddingSharder[EmbeddingTowerCollection]):
    def __init__(self, fused_params=None, qcomm_codecs_registry=None) -> None:
        super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._tower_sharder = EmbeddingTowerSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)

    def shard(self, module: EmbeddingTowerCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedEmbeddingTowerCollection:
        device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedEmbeddingTowerCollection(module, params, self._tower_sharder._ebc_sharder, self._tower_sharder._ec_sharder, env, device)

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return self._tower_sharder.sharding_types(compute_device_type)

    def shardable_parameters(self, module: EmbeddingTowerCollection) -> Dict[str, nn.Parameter]:
        out: Dict[str, nn.Parameter] = {}
        for tower in module.towers:
            out.update(self._tower_sharder.shardable_parameters(tower))
        return out

    @property
    def module_type(self) -> Type[EmbeddingTowerCollection]:
        return EmbeddingTowerCollection
