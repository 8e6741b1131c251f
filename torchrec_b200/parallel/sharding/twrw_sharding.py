"""Table-row-wise sharding: a table lives on ONE host, its rows range-split over that host's ranks.

Reference: ``torchrec/distributed/sharding/twrw_sharding.py`` - ``BaseTwRwEmbeddingSharding`` :68-277, ``TwRwSparseFeaturesDist`` :280-432,
``TwRwPooledEmbeddingDist`` :435-648, ``TwRwPooledEmbeddingSharding`` :651-714.
Input: ids bucketized by row owner INSIDE the host (``local_size`` buckets), keys laid out rank-major, one KJT all-to-all with ``features_per_rank``.
Output, two hops that match the fabric: (1) reduce-scatter inside the host over NVLink - sums the row partials and leaves each local rank with the
samples of its cross-host peers, (2) pooled all-to-all across hosts between same-local-rank peers - one rail per GPU.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..comm import get_local_size, intra_and_cross_node_pg, intra_and_cross_node_pg_2D
from ..dist_data import KJTAllToAll, PooledEmbeddingsAllToAll, PooledEmbeddingsReduceScatter
from ..embedding_sharding import (
    BaseEmbeddingDist,
    BaseSparseFeaturesDist,
    C,
    EmbeddingShardingContext,
    EmbeddingShardingInfo,
    F,
    T,
    W,
    bucketize_kjt_before_all2all,
)
from ..embedding_types import BaseEmbeddingLookup, ShardedEmbeddingTable
from ..types import Awaitable, CommOp, QuantizedCommCodecs, ShardingEnv, ShardingEnv2D, ShardMetadata
from .common import BaseShardingCommon, make_shard_table, rank_of, shards_of_info
from .tw_sharding import _global_md


class BaseTwRwEmbeddingSharding(BaseShardingCommon[C, F, T, W]):
    def __init__(self, sharding_infos: List[EmbeddingShardingInfo], env: ShardingEnv, device: Optional[torch.device] = None, need_pos: bool = False,
                 qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        self._local_size: int = getattr(env, "node_group_size", None) or get_local_size(env.world_size)
        assert env.world_size % self._local_size == 0, f"world size {env.world_size} is not a multiple of the host size {self._local_size}"
        super().__init__(sharding_infos, env, device, need_pos, qcomm_codecs_registry)
        if isinstance(env, ShardingEnv2D):
            self._intra_pg, self._cross_pg = intra_and_cross_node_pg_2D(env, device=self._device)
        else:
            self._intra_pg, self._cross_pg = intra_and_cross_node_pg(device=self._device, backend=dist.get_backend(self._pg))

    def _node_of(self, rank: int) -> int:
        return rank // self._local_size

    def _shard(self, sharding_infos: List[EmbeddingShardingInfo]) -> List[List[ShardedEmbeddingTable]]:
        tables_per_rank: List[List[ShardedEmbeddingTable]] = [[] for _ in range(self._world_size)]
        L = self._local_size
        for info in sharding_infos:
            cfg = info.embedding_config
            gmd = _global_md(info)
            # column blocks (grid) -> per block the row shards of one host; plain TWRW has a single block
            blocks: Dict[int, List[ShardMetadata]] = {}
            for s in shards_of_info(info):
                blocks.setdefault(s.shard_offsets[1], []).append(s)
            for col0 in sorted(blocks):
                shards = blocks[col0]
                node = self._node_of(rank_of(shards[0].placement))
                by_rank = {rank_of(s.placement): s for s in shards}
                for rank in range(node * L, (node + 1) * L):
                    s = by_rank.get(rank)
                    if s is None:
                        s = ShardMetadata(shard_offsets=[cfg.num_embeddings, col0], shard_sizes=[0, shards[0].shard_sizes[1]], placement=f"rank:{rank}/{self._device.type}")
                    tables_per_rank[rank].append(make_shard_table(info, s, s.shard_sizes[0], s.shard_sizes[1], gmd))
        return tables_per_rank

    # per-host views: local rank 0 of each host is its representative (all ranks of a host carry the same features)
    def _reps(self) -> List[int]:
        return list(range(0, self._world_size, self._local_size))

    def embedding_dims(self) -> List[int]:
        return [d for r in self._reps() for g in self._grouped_embedding_configs_per_rank[r] for d in g.embedding_dims()]

    def embedding_names(self) -> List[str]:
        return [n for r in self._reps() for g in self._grouped_embedding_configs_per_rank[r] for n in g.embedding_names()]

    def embedding_names_per_rank(self) -> List[List[str]]:
        raise NotImplementedError

    def embedding_shard_metadata(self) -> List[Optional[ShardMetadata]]:
        return [t.local_metadata for r in self._reps() for g in self._grouped_embedding_configs_per_rank[r] for t in g.embedding_tables for _ in t.feature_names]

    def feature_names(self) -> List[str]:
        return [f for r in self._reps() for g in self._grouped_embedding_configs_per_rank[r] for f in g.feature_names()]

    def embedding_tables(self) -> List[ShardedEmbeddingTable]:
        return [t for r in self._reps() for g in self._grouped_embedding_configs_per_rank[r] for t in g.embedding_tables]

    def _features_per_node(self) -> List[int]:
        return [sum(g.num_features() for g in self._grouped_embedding_configs_per_rank[r]) for r in self._reps()]

    def _dim_sum_per_node(self) -> List[int]:
        return [sum(g.dim_sum() for g in self._grouped_embedding_configs_per_rank[r]) for r in self._reps()]

    def _emb_dim_per_node_per_feature(self) -> List[List[int]]:
        return [[d for g in self._grouped_embedding_configs_per_rank[r] for d in g.embedding_dims()] for r in self._reps()]

    def _get_feature_hash_sizes(self) -> List[int]:
        return [h for r in self._reps() for g in self._grouped_embedding_configs_per_rank[r] for h in g.feature_hash_sizes()]

    def _row_boundaries(self) -> Optional[List[List[int]]]:
        """Per feature (host-major order) the row offsets of its ``local_size`` shards + the table size, or None when every table is split evenly."""
        L = self._local_size
        out: List[List[int]] = []
        even = True
        for rep in self._reps():
            per_rank = [[t for g in self._grouped_embedding_configs_per_rank[rep + l] for t in g.embedding_tables] for l in range(L)]
            for ti, t0 in enumerate(per_rank[0]):
                block = -(-t0.num_embeddings // L)
                offs = [per_rank[l][ti].local_metadata.shard_offsets[0] for l in range(L)]
                rows = [per_rank[l][ti].local_rows for l in range(L)]
                if rows != [max(0, min(block, t0.num_embeddings - l * block)) for l in range(L)]:
                    even = False
                for _ in t0.feature_names:
                    out.append(offs + [t0.num_embeddings])
        return None if even else out


class TwRwSparseFeaturesDist(BaseSparseFeaturesDist[KeyedJaggedTensor]):
    def __init__(self, pg: dist.ProcessGroup, local_size: int, features_per_rank: List[int], feature_hash_sizes: List[int], device: Optional[torch.device] = None,
                 has_feature_processor: bool = False, need_pos: bool = False, keep_original_indices: bool = False,
                 embedding_shard_metadata: Optional[List[List[int]]] = None, is_sequence: bool = False) -> None:
        super().__init__()
        W_ = pg.size()
        assert W_ % local_size == 0
        self._world_size, self._local_size = W_, local_size
        self._num_cross_nodes = W_ // local_size
        per_node = [features_per_rank[n * local_size] for n in range(self._num_cross_nodes)]
        F_total = sum(per_node)
        assert F_total == len(feature_hash_sizes)
        self.register_buffer("_feature_block_size_tensor", torch.tensor([-(-h // local_size) for h in feature_hash_sizes], device=device, dtype=torch.int64), persistent=False)
        self._row_pos = [torch.tensor(b, device=device, dtype=torch.int64) for b in embedding_shard_metadata] if embedding_shard_metadata else None
        # bucketized keys are [bucket l][feature f]; ranks want [(host n, local l)][features of host n]
        perm, start = [], 0
        for n in range(self._num_cross_nodes):
            for l in range(local_size):
                perm.extend(l * F_total + f for f in range(start, start + per_node[n]))
            start += per_node[n]
        self._permute = perm
        self.register_buffer("_permute_tensor", torch.tensor(perm, device=device, dtype=torch.int64), persistent=False)
        self._dist = KJTAllToAll(pg=pg, splits=features_per_rank)
        self._has_feature_processor, self._need_pos, self._keep_original_indices, self._is_sequence = has_feature_processor, need_pos, keep_original_indices, is_sequence
        self.unbucketize_permute_tensor: Optional[torch.Tensor] = None

    def forward(self, sparse_features: KeyedJaggedTensor) -> Awaitable[Awaitable[KeyedJaggedTensor]]:
        bucketized, self.unbucketize_permute_tensor = bucketize_kjt_before_all2all(
            sparse_features, num_buckets=self._local_size, block_sizes=self._feature_block_size_tensor, output_permute=self._is_sequence,
            bucketize_pos=(self._has_feature_processor if sparse_features.weights_or_none() is None else self._need_pos),
            block_bucketize_row_pos=self._row_pos, keep_original_indices=self._keep_original_indices)
        return self._dist(bucketized.permute(self._permute, self._permute_tensor))


class TwRwPooledEmbeddingDist(BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]):
    def __init__(self, rank: int, cross_pg: dist.ProcessGroup, intra_pg: dist.ProcessGroup, dim_sum_per_node: List[int], emb_dim_per_node_per_feature: List[List[int]],
                 device: Optional[torch.device] = None, qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None, callbacks: Optional[List[Any]] = None) -> None:
        super().__init__()
        reg = qcomm_codecs_registry or {}
        self._rank = rank
        self._intra_pg, self._cross_pg = intra_pg, cross_pg
        self._local_size = intra_pg.size()
        self._num_nodes = cross_pg.size()
        self._intra_dist = PooledEmbeddingsReduceScatter(intra_pg, reg.get(CommOp.POOLED_EMBEDDINGS_REDUCE_SCATTER.name))
        self._cross_dist = PooledEmbeddingsAllToAll(cross_pg, dim_sum_per_node, device, callbacks, reg.get(CommOp.POOLED_EMBEDDINGS_ALL_TO_ALL.name))
        self._emb_dim_per_node_per_feature = emb_dim_per_node_per_feature

    def forward(self, local_embs: torch.Tensor, sharding_ctx: Optional[EmbeddingShardingContext] = None) -> Awaitable[torch.Tensor]:
        if sharding_ctx is not None and sharding_ctx.variable_batch_per_feature:
            raise NotImplementedError("variable batch per feature is served by the engine path (ShardedEmbeddingBagCollection), not by the composable TWRW dist")
        N, L = self._num_nodes, self._local_size
        B_global, D = local_embs.shape
        bpr = sharding_ctx.batch_size_per_rank if sharding_ctx is not None and sharding_ctx.batch_size_per_rank else None
        if bpr is not None and len(set(bpr)) > 1:
            # uneven batches: regroup the sample blocks by local rank, reduce-scatter-v inside the host, all-to-all-v across hosts
            blocks = torch.split(local_embs, bpr, dim=0)
            order = [n * L + l for l in range(L) for n in range(N)]
            x = torch.cat([blocks[r] for r in order], dim=0)
            intra_splits = [sum(bpr[n * L + l] for n in range(N)) for l in range(L)]
            my_l = self._rank % L
            rs = self._intra_dist(x, input_splits=intra_splits).wait()
            return self._cross_dist(rs, batch_size_per_rank=[bpr[n * L + my_l] for n in range(N)])
        B = B_global // (N * L)
        x = local_embs.view(N, L, B, D).transpose(0, 1).reshape(L * N * B, D)
        rs = self._intra_dist(x).wait()
        return self._cross_dist(rs)


class TwRwPooledEmbeddingSharding(BaseTwRwEmbeddingSharding[EmbeddingShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[KeyedJaggedTensor]:
        assert self._pg is not None
        return TwRwSparseFeaturesDist(self._pg, self._local_size, self.features_per_rank(), self._get_feature_hash_sizes(), device if device is not None else self._device,
                                      has_feature_processor=any(g.has_feature_processor for g in self._grouped_embedding_configs), need_pos=self._need_pos,
                                      embedding_shard_metadata=self._row_boundaries())

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return self._pooled_lookup(device, fused_params, feature_processor)

    def _output_callbacks(self) -> Optional[List[Any]]:
        return None

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]:
        assert self._intra_pg is not None and self._cross_pg is not None
        return TwRwPooledEmbeddingDist(self._rank, self._cross_pg, self._intra_pg, self._dim_sum_per_node(), self._emb_dim_per_node_per_feature(),
                                       device if device is not None else self._device, self.qcomm_codecs_registry, callbacks=self._output_callbacks())
