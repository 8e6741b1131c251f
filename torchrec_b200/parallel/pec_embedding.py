"""Sharded Prioritized-Embedding-Communication collection (reference torchrec/distributed/pec_embedding.py:49-495).

Composition around ``ShardedEmbeddingCollection``. On the owner side of the input dist every batch's ids are partitioned against the
previous batch's ids (``pec_collision_handlers``): the non-overlapped partition touches rows the in-flight step is NOT updating, so its
lookup may run early (``prefetch_nonoverlapped`` - called by a pipeline for batch i+1 before batch i's backward); the overlapped
partition is looked up at forward time, after that update. Both partial results are merged on the owner with one ``index_select``
(autograd re-splits the gradient) and leave through the ordinary sequence output dist. A row id lives in exactly one partition, so the
two fused backward passes update disjoint rows: results are bit-identical to the plain sharded collection (tests/test_extra_sharders_gloo).

Difference to the reference: its two partitions travel back in two all-to-alls (overlapped first); here the merge happens before one
output dist - on an NVSwitch node the win is the early HBM gather, the dist itself is a few microseconds of peer stores."""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Type

import torch
from torch import nn

from ..modules.pec_embedding_modules import PECEmbeddingCollection
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .embedding import EmbeddingCollectionContext, EmbeddingCollectionSharder
from .embedding_types import BaseEmbeddingSharder, KJTList
from .pec_collision_handlers import CollisionResult, create_collision_handler, split_features_by_values_mask
from .types import LazyAwaitable, ParameterSharding, ShardedModule, ShardingEnv


class PECEmbeddingCollectionContext(EmbeddingCollectionContext):
    def __init__(self) -> None:
        super().__init__()
        self.prev_remapped_feature_values: Optional[torch.Tensor] = None  # row keys of the previous batch (set by the module / pipeline)
        self.collision: Optional[CollisionResult] = None
        self.nol_prefetched: Optional[tuple] = None                       # (ol_kjt, nol_kjt, permutation, nol_rows) from prefetch_nonoverlapped


class ShardedPECEmbeddingCollection(ShardedModule):
    def __init__(self, module: PECEmbeddingCollection, params: Dict[str, ParameterSharding], ec_sharder: EmbeddingCollectionSharder, env: ShardingEnv,
                 device: torch.device) -> None:
        super().__init__()
        self._env, self._device = env, device
        self._embedding_collection = ec_sharder.shard(module._embedding_collection, params, env=env, device=device)
        eng = self._embedding_collection._engine
        shard_base: Dict[int, int] = {}
        total = 0
        for u in eng.local_units:  # units of features sharing a table read the same shard: same key range
            if u.shard.local_idx not in shard_base:
                shard_base[u.shard.local_idx] = total
                total += u.shard.rows
        self._handler = create_collision_handler(module._checker_type, [shard_base[u.shard.local_idx] for u in eng.local_units], total, device)
        self._last_keys: Optional[torch.Tensor] = None
        self.stats = {"values": 0, "overlapped": 0}

    # -- ShardedModule phases ---------------------------------------------------------------------------------------------------------
    def create_context(self) -> PECEmbeddingCollectionContext:
        ctx = PECEmbeddingCollectionContext()
        ctx.prev_remapped_feature_values = self._last_keys
        return ctx

    def input_dist(self, ctx: PECEmbeddingCollectionContext, features: KeyedJaggedTensor):
        return self._embedding_collection.input_dist(ctx, features)

    def detect_collisions(self, ctx: PECEmbeddingCollectionContext, dist_input: KJTList) -> List[CollisionResult]:
        res = self._handler.detect_collisions(dist_input[0], ctx.prev_remapped_feature_values)
        ctx.collision = res
        return [res]

    @torch.no_grad()
    def _partition(self, ctx: PECEmbeddingCollectionContext, dist_input: KJTList):
        if ctx.collision is None:
            self.detect_collisions(ctx, dist_input)
        assert ctx.collision is not None
        return split_features_by_values_mask(dist_input[0], ctx.collision.forward_overlap_mask)

    def prefetch_nonoverlapped(self, ctx: PECEmbeddingCollectionContext, dist_input: KJTList) -> None:
        """Early stage for the NEXT batch: partition it against the batch in flight and gather the rows that batch is not updating."""
        ol_kjt, nol_kjt, perm = self._partition(ctx, dist_input)
        ctx.nol_prefetched = (ol_kjt, nol_kjt, perm, self._embedding_collection._engine.lookup(nol_kjt))

    def compute(self, ctx: PECEmbeddingCollectionContext, dist_input: KJTList):
        eng = self._embedding_collection._engine
        if ctx.nol_prefetched is not None:
            ol_kjt, nol_kjt, perm, nol_rows = ctx.nol_prefetched
            ctx.nol_prefetched = None
        else:
            ol_kjt, nol_kjt, perm = self._partition(ctx, dist_input)
            nol_rows = eng.lookup(nol_kjt)
        ol_rows = eng.lookup(ol_kjt)                       # rows the previous step updated: looked up after its backward
        merged = torch.cat([ol_rows, nol_rows], dim=0).index_select(0, perm.forward_permute)
        assert ctx.collision is not None
        self._last_keys = ctx.collision.remapped_feature_values
        self.stats["values"] += int(perm.forward_permute.numel())
        self.stats["overlapped"] += int(perm.num_overlapped)
        return [merged], dist_input[0]

    def output_dist(self, ctx: PECEmbeddingCollectionContext, output) -> LazyAwaitable[Dict[str, JaggedTensor]]:
        return self._embedding_collection.output_dist(ctx, output)

    def compute_and_output_dist(self, ctx: PECEmbeddingCollectionContext, input: KJTList) -> LazyAwaitable[Dict[str, JaggedTensor]]:
        return self.output_dist(ctx, self.compute(ctx, input))

    # -- parameters / state -------------------------------------------------------------------------------------------------------------
    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        p = prefix + "." if prefix else ""
        yield from self._embedding_collection.sharded_parameter_names(p + "_embedding_collection")

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
        return self._embedding_collection.fused_optimizer


class PECEmbeddingCollectionSharder(BaseEmbeddingSharder[PECEmbeddingCollection]):
    def __init__(self, ec_sharder: Optional[EmbeddingCollectionSharder] = None, fused_params=None, qcomm_codecs_registry=None) -> None:
        super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._ec_sharder = ec_sharder or EmbeddingCollectionSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)

    def shard(self, module: PECEmbeddingCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedPECEmbeddingCollection:
        device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedPECEmbeddingCollection(module, params, self._ec_sharder, env, device)

    def shardable_parameters(self, module: PECEmbeddingCollection) -> Dict[str, nn.Parameter]:
        return self._ec_sharder.shardable_parameters(module._embedding_collection)

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return self._ec_sharder.sharding_types(compute_device_type)

    @property
    def module_type(self) -> Type[PECEmbeddingCollection]:
        return PECEmbeddingCollection
