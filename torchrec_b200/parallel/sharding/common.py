"""Bookkeeping shared by the per-type ``EmbeddingSharding`` classes of this package.

Every type answers ONE question differently - "which rectangles of which tables live on which rank" (``_shard``) - and picks its collectives;
everything else (grouping the local rectangles into kernels, naming, dims, lookups) is the same and lives here. The rectangles come from the plan's
``ParameterSharding.sharding_spec`` exactly like ``engine.shards_of`` reads them.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Generic, List, Optional, Tuple, TypeVar

import torch
import torch.distributed as dist
from torch import nn

from ...modules.embedding_configs import DataType
from ..embedding_lookup import GroupedEmbeddingsLookup, GroupedPooledEmbeddingsLookup
from ..embedding_sharding import C, EmbeddingSharding, EmbeddingShardingInfo, F, T, W, group_tables
from ..embedding_types import BaseEmbeddingLookup, EmbeddingComputeKernel, GroupedEmbeddingConfig, ShardedEmbeddingTable
from ..types import QuantizedCommCodecs, ShardingEnv, ShardMetadata


def rank_of(placement: Any) -> int:
    """``rank:3/cuda:3`` (string or torch ``_remote_device``) -> 3."""
    if hasattr(placement, "rank") and callable(placement.rank):
        r = placement.rank()
        if r is not None:
            return int(r)
    s = str(placement)
    return int(s.split("/")[0].split(":")[1]) if s.startswith("rank:") else 0


def make_shard_table(info: EmbeddingShardingInfo, shard: Optional[ShardMetadata], rows: int, cols: int, global_metadata: Any = None) -> ShardedEmbeddingTable:
    cfg = info.embedding_config
    kernel = info.param_sharding.compute_kernel
    return ShardedEmbeddingTable(
        num_embeddings=cfg.num_embeddings, embedding_dim=cfg.embedding_dim, name=cfg.name, data_type=cfg.data_type, feature_names=list(cfg.feature_names),
        weight_init_max=cfg.weight_init_max, weight_init_min=cfg.weight_init_min, init_fn=getattr(cfg, "init_fn", None), need_pos=getattr(cfg, "need_pos", False),
        total_num_buckets=getattr(cfg, "total_num_buckets", None), use_virtual_table=getattr(cfg, "use_virtual_table", False),
        enable_embedding_update=getattr(cfg, "enable_embedding_update", False),
        pooling=getattr(cfg, "pooling", None) or ShardedEmbeddingTable.__dataclass_fields__["pooling"].default,
        is_weighted=getattr(cfg, "is_weighted", False), has_feature_processor=getattr(cfg, "has_feature_processor", False),
        embedding_names=list(getattr(cfg, "embedding_names", None) or cfg.feature_names),
        compute_kernel=kernel if isinstance(kernel, EmbeddingComputeKernel) else EmbeddingComputeKernel(str(kernel)),
        local_rows=rows, local_cols=cols, local_metadata=shard, global_metadata=global_metadata, fused_params=info.fused_params,
    )


def shards_of_info(info: EmbeddingShardingInfo) -> List[ShardMetadata]:
    spec = info.param_sharding.sharding_spec
    assert spec is not None, f"table {info.embedding_config.name}: plan entry has no sharding_spec"
    return list(spec.shards)  # type: ignore[attr-defined]


class BaseShardingCommon(EmbeddingSharding[C, F, T, W]):
    """Env / device / grouped configs / names shared by all types. Subclasses implement ``_shard`` and the three ``create_*``."""

    def __init__(self, sharding_infos: List[EmbeddingShardingInfo], env: ShardingEnv, device: Optional[torch.device] = None,
                 need_pos: bool = False, qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__(qcomm_codecs_registry=qcomm_codecs_registry)
        self._env = env
        self._pg: Optional[dist.ProcessGroup] = env.process_group
        self._world_size: int = env.world_size
        self._rank: int = env.rank
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        self._need_pos = need_pos
        self._sharding_infos = sharding_infos
        self._sharded_tables_per_rank: List[List[ShardedEmbeddingTable]] = self._shard(sharding_infos)
        self._grouped_embedding_configs_per_rank: List[List[GroupedEmbeddingConfig]] = group_tables(self._sharded_tables_per_rank)
        self._grouped_embedding_configs: List[GroupedEmbeddingConfig] = self._grouped_embedding_configs_per_rank[self._rank] if self._rank < len(self._grouped_embedding_configs_per_rank) else []
        self._copy_weights()

    def _shard(self, sharding_infos: List[EmbeddingShardingInfo]) -> List[List[ShardedEmbeddingTable]]:
        raise NotImplementedError

    # -- initial values: the local rectangles of the unsharded parameters (``EmbeddingShardingInfo.param``) --
    def _copy_weights(self) -> None:
        self._init_rows: Dict[Tuple[str, int, int], torch.Tensor] = {}
        by_name = {i.embedding_config.name: i for i in self._sharding_infos}
        for g in self._grouped_embedding_configs:
            for t in g.embedding_tables:
                p = by_name[t.name].param
                if p is None or getattr(p, "device", torch.device("meta")).type == "meta":
                    continue
                md = t.local_metadata
                r0, c0 = (md.shard_offsets[0], md.shard_offsets[1]) if md is not None else (0, 0)
                self._init_rows[(t.name, r0, c0)] = p.detach()[r0 : r0 + t.local_rows, c0 : c0 + t.local_cols]

    def _load_initial_weights(self, lookup: nn.Module) -> None:
        with torch.no_grad():
            for m in getattr(lookup, "_emb_modules", []):
                for t, w in zip(m.config.embedding_tables, m.split_embedding_weights()):
                    md = t.local_metadata
                    key = (t.name, md.shard_offsets[0] if md is not None else 0, md.shard_offsets[1] if md is not None else 0)
                    src = self._init_rows.get(key)
                    if src is not None:
                        w.copy_(src.to(w.dtype))
                inner = getattr(m, "emb_module", None)
                if inner is not None and hasattr(inner, "load_rows_changed"):
                    inner.load_rows_changed()

    # -- lookups --
    def _codec(self, op_name: str) -> Optional[QuantizedCommCodecs]:
        return (self._qcomm_codecs_registry or {}).get(op_name)

    def _with_fused_params(self, fused_params: Optional[Dict[str, Any]]) -> List[GroupedEmbeddingConfig]:
        if not fused_params:
            return self._grouped_embedding_configs
        out = []
        for g in self._grouped_embedding_configs:
            g = copy.copy(g)
            g.fused_params = {**(fused_params or {}), **(g.fused_params or {})}
            out.append(g)
        return out

    def _pooled_lookup(self, device: Optional[torch.device], fused_params: Optional[Dict[str, Any]], feature_processor: Optional[nn.Module],
                       pg: Optional[dist.ProcessGroup] = None, scale_weight_gradients: bool = True) -> GroupedPooledEmbeddingsLookup:
        lk = GroupedPooledEmbeddingsLookup(self._with_fused_params(fused_params), pg if pg is not None else self._pg, device if device is not None else self._device,
                                           feature_processor, scale_weight_gradients, env=self._env)
        self._load_initial_weights(lk)
        return lk

    def _sequence_lookup(self, device: Optional[torch.device], fused_params: Optional[Dict[str, Any]], pg: Optional[dist.ProcessGroup] = None) -> GroupedEmbeddingsLookup:
        lk = GroupedEmbeddingsLookup(self._with_fused_params(fused_params), pg if pg is not None else self._pg, device if device is not None else self._device, env=self._env)
        self._load_initial_weights(lk)
        return lk

    # -- naming: rank-major, group order inside a rank (the order the lookups emit columns) --
    def _tables_of_rank(self, rank: int) -> List[ShardedEmbeddingTable]:
        return [t for g in self._grouped_embedding_configs_per_rank[rank] for t in g.embedding_tables]

    def embedding_tables(self) -> List[ShardedEmbeddingTable]:
        return [t for r in range(len(self._grouped_embedding_configs_per_rank)) for t in self._tables_of_rank(r)]

    def embedding_dims(self) -> List[int]:
        return [d for g_rank in self._grouped_embedding_configs_per_rank for g in g_rank for d in g.embedding_dims()]

    def embedding_names(self) -> List[str]:
        return [n for g_rank in self._grouped_embedding_configs_per_rank for g in g_rank for n in g.embedding_names()]

    def embedding_names_per_rank(self) -> List[List[str]]:
        return [[n for g in g_rank for n in g.embedding_names()] for g_rank in self._grouped_embedding_configs_per_rank]

    def embedding_shard_metadata(self) -> List[Optional[ShardMetadata]]:
        return [t.local_metadata for g_rank in self._grouped_embedding_configs_per_rank for g in g_rank for t in g.embedding_tables for _ in t.feature_names]

    def feature_names(self) -> List[str]:
        return [f for g_rank in self._grouped_embedding_configs_per_rank for g in g_rank for f in g.feature_names()]

    def feature_names_per_rank(self) -> List[List[str]]:
        return [[f for g in g_rank for f in g.feature_names()] for g_rank in self._grouped_embedding_configs_per_rank]

    def features_per_rank(self) -> List[int]:
        return [sum(g.num_features() for g in g_rank) for g_rank in self._grouped_embedding_configs_per_rank]

    def _dim_sum_per_rank(self) -> List[int]:
        return [sum(g.dim_sum() for g in g_rank) for g_rank in self._grouped_embedding_configs_per_rank]

    def _emb_dim_per_rank_per_feature(self) -> List[List[int]]:
        return [[d for g in g_rank for d in g.embedding_dims()] for g_rank in self._grouped_embedding_configs_per_rank]
