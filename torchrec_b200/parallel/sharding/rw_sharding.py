"""Row-wise sharding: a table's rows range-split over the ranks.

Reference: ``torchrec/distributed/sharding/rw_sharding.py`` - ``get_embedding_shard_metadata`` :73-107, ``BaseRwEmbeddingSharding`` :110-330,
``RwSparseFeaturesDist`` :333-514, ``RwPooledEmbeddingDist`` :517-601, ``RwPooledEmbeddingSharding`` :644-702, ``RwSparseFeaturesWriteDist`` :705-812,
inference variants :604-641 / :834-1007.
Input: ids bucketized by row range (``block_size = ceil(rows / W)``, or the plan's uneven boundaries), then a KJT all-to-all with ``[F] * W`` splits.
Output: reduce-scatter of the per-rank partial pools ``[B_global, D] -> [B_local, D]`` (variable batch: reduce-scatter-v).
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..dist_data import (
    EmbeddingsAllToOneReduce,
    KJTAllToAll,
    KJTOneToAll,
    PooledEmbeddingsReduceScatter,
    VariableBatchPooledEmbeddingsReduceScatter,
)
from ..embedding_lookup import GroupedEmbeddingsUpdate, InferGroupedPooledEmbeddingsLookup
from ..embedding_sharding import (
    BaseEmbeddingDist,
    BaseSparseFeaturesDist,
    BaseSparseFeaturesWriteDist,
    C,
    EmbeddingShardingContext,
    EmbeddingShardingInfo,
    F,
    T,
    W,
    bucketize_embeddings_before_all2all_write,
    bucketize_kjt_before_all2all,
    bucketize_kjt_inference,
)
from ..embedding_types import BaseEmbeddingLookup, GroupedEmbeddingConfig, InputDistOutputs, KJTList, ShardedEmbeddingTable
from ..types import Awaitable, CommOp, NoWait, NullShardingContext, QuantizedCommCodecs, ShardingEnv, ShardMetadata
from .common import BaseShardingCommon, make_shard_table, rank_of, shards_of_info
from .tw_sharding import _global_md


def get_embedding_shard_metadata(grouped_embedding_configs_per_rank: List[List[GroupedEmbeddingConfig]]) -> Tuple[List[List[int]], bool]:
    """Per feature the row offsets of its shards over the ranks, and whether any table is split unevenly (reference :73-107)."""
    is_even = True
    world = len(grouped_embedding_configs_per_rank)
    offsets: List[List[int]] = []
    if world == 0 or not grouped_embedding_configs_per_rank[0]:
        return offsets, is_even
    per_rank = [[t for g in gs for t in g.embedding_tables] for gs in grouped_embedding_configs_per_rank]
    for ti, t0 in enumerate(per_rank[0]):
        block = math.ceil(t0.num_embeddings / world)
        offs = [per_rank[r][ti].local_metadata.shard_offsets[0] if per_rank[r][ti].local_metadata is not None else min(r * block, t0.num_embeddings) for r in range(world)]
        rows = [per_rank[r][ti].local_rows for r in range(world)]
        expect = [max(0, min(block, t0.num_embeddings - r * block)) for r in range(world)]
        if rows != expect:
            is_even = False
        for _ in t0.feature_names:
            offsets.append(offs + [t0.num_embeddings])
    return offsets, is_even


def get_block_sizes_runtime_device(block_sizes: List[int], runtime_device: torch.device, tensor_cache: Dict[str, Tuple[torch.Tensor, List[torch.Tensor]]],
                                   embedding_shard_metadata: Optional[List[List[int]]] = None, dtype: torch.dtype = torch.int32) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Block sizes (and uneven boundaries) as tensors on the device the KJT lives on, cached per device (reference inference path)."""
    key = f"{runtime_device}/{dtype}"
    if key not in tensor_cache:
        tensor_cache[key] = (torch.tensor(block_sizes, device=runtime_device, dtype=dtype),
                             [torch.tensor(r, device=runtime_device, dtype=dtype) for r in (embedding_shard_metadata or [])])
    return tensor_cache[key]


class BaseRwEmbeddingSharding(BaseShardingCommon[C, F, T, W]):
    def _shard(self, sharding_infos: List[EmbeddingShardingInfo]) -> List[List[ShardedEmbeddingTable]]:
        tables_per_rank: List[List[ShardedEmbeddingTable]] = [[] for _ in range(self._world_size)]
        for info in sharding_infos:
            shards = shards_of_info(info)
            gmd = _global_md(info)
            by_rank = {rank_of(s.placement): s for s in shards}
            cfg = info.embedding_config
            for rank in range(self._world_size):
                s = by_rank.get(rank)
                if s is None:  # the plan gives this rank no rows: an empty shard keeps the feature layout identical on every rank
                    s = ShardMetadata(shard_offsets=[cfg.num_embeddings, 0], shard_sizes=[0, cfg.embedding_dim], placement=f"rank:{rank}/{self._device.type}")
                tables_per_rank[rank].append(make_shard_table(info, s, s.shard_sizes[0], s.shard_sizes[1], gmd))
        return tables_per_rank

    # every rank looks up every feature: names / dims are those of ONE rank
    def embedding_dims(self) -> List[int]:
        return [d for g in self._grouped_embedding_configs for d in g.embedding_dims()]

    def embedding_names(self) -> List[str]:
        return [n for g in self._grouped_embedding_configs for n in g.embedding_names()]

    def embedding_names_per_rank(self) -> List[List[str]]:
        raise NotImplementedError("row-wise shards carry every feature on every rank")

    def embedding_shard_metadata(self) -> List[Optional[ShardMetadata]]:
        return [t.local_metadata for g in self._grouped_embedding_configs for t in g.embedding_tables for _ in t.feature_names]

    def embedding_tables(self) -> List[ShardedEmbeddingTable]:
        return [t for g in self._grouped_embedding_configs for t in g.embedding_tables]

    def feature_names(self) -> List[str]:
        return [f for g in self._grouped_embedding_configs for f in g.feature_names()]

    def _get_num_features(self) -> int:
        return sum(g.num_features() for g in self._grouped_embedding_configs)

    def _get_feature_hash_sizes(self) -> List[int]:
        return [h for g in self._grouped_embedding_configs for h in g.feature_hash_sizes()]

    def _get_feature_total_num_buckets(self) -> Optional[List[int]]:
        out = [getattr(t, "total_num_buckets", None) for g in self._grouped_embedding_configs for t in g.embedding_tables for _ in t.feature_names]
        return None if any(b is None for b in out) or not out else [int(b) for b in out]

    def _row_boundaries(self) -> Optional[List[List[int]]]:
        offs, even = get_embedding_shard_metadata(self._grouped_embedding_configs_per_rank)
        return None if even else offs


class RwSparseFeaturesDist(BaseSparseFeaturesDist[KeyedJaggedTensor]):
    """Bucketize by row owner, then all-to-all with equal key splits (every rank owns a slice of every table)."""

    def __init__(self, pg: dist.ProcessGroup, num_features: int, feature_hash_sizes: List[int], feature_total_num_buckets: Optional[List[int]] = None,
                 device: Optional[torch.device] = None, is_sequence: bool = False, has_feature_processor: bool = False, need_pos: bool = False,
                 keep_original_indices: bool = False, embedding_shard_metadata: Optional[List[List[int]]] = None) -> None:
        super().__init__()
        self._world_size: int = pg.size()
        self._num_features = num_features
        if feature_total_num_buckets is not None:
            # bucket-aware (ZCH) tables: whole buckets go to a rank
            block = [(-(-b // self._world_size)) * (-(-h // b)) for h, b in zip(feature_hash_sizes, feature_total_num_buckets)]
        else:
            block = [-(-h // self._world_size) for h in feature_hash_sizes]
        self.register_buffer("_feature_block_size_tensor", torch.tensor(block, device=device, dtype=torch.int64), persistent=False)
        self._feature_total_num_blocks = None if feature_total_num_buckets is None else torch.tensor(feature_total_num_buckets, device=device, dtype=torch.int64)
        self._row_pos: Optional[List[torch.Tensor]] = None
        if embedding_shard_metadata:
            self._row_pos = [torch.tensor(b, device=device, dtype=torch.int64) for b in embedding_shard_metadata]
        self._dist = KJTAllToAll(pg=pg, splits=[num_features] * self._world_size)
        self._is_sequence = is_sequence
        self._has_feature_processor = has_feature_processor
        self._need_pos = need_pos
        self._keep_original_indices = keep_original_indices
        self.unbucketize_permute_tensor: Optional[torch.Tensor] = None

    def forward(self, sparse_features: KeyedJaggedTensor) -> Awaitable[Awaitable[KeyedJaggedTensor]]:
        bucketized, self.unbucketize_permute_tensor = bucketize_kjt_before_all2all(
            sparse_features, num_buckets=self._world_size, block_sizes=self._feature_block_size_tensor, total_num_blocks=self._feature_total_num_blocks,
            output_permute=self._is_sequence, bucketize_pos=(self._has_feature_processor if sparse_features.weights_or_none() is None else self._need_pos),
            block_bucketize_row_pos=self._row_pos, keep_original_indices=self._keep_original_indices,
        )
        return self._dist(bucketized)


class RwPooledEmbeddingDist(BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]):
    def __init__(self, pg: dist.ProcessGroup, embedding_dims: List[int], qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__()
        codecs = (qcomm_codecs_registry or {}).get(CommOp.POOLED_EMBEDDINGS_REDUCE_SCATTER.name)
        self._dist = PooledEmbeddingsReduceScatter(pg, codecs)
        self._variable_dist = VariableBatchPooledEmbeddingsReduceScatter(pg, codecs)
        self._embedding_dims = embedding_dims

    def forward(self, local_embs: torch.Tensor, sharding_ctx: Optional[EmbeddingShardingContext] = None) -> Awaitable[torch.Tensor]:
        if sharding_ctx is None:
            return self._dist(local_embs)
        if sharding_ctx.variable_batch_per_feature:
            return self._variable_dist(local_embs, batch_size_per_rank_per_feature=sharding_ctx.batch_size_per_rank_per_feature, embedding_dims=self._embedding_dims)
        return self._dist(local_embs, input_splits=sharding_ctx.batch_size_per_rank or None)


class RwPooledEmbeddingSharding(BaseRwEmbeddingSharding[EmbeddingShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[KeyedJaggedTensor]:
        assert self._pg is not None
        return RwSparseFeaturesDist(self._pg, self._get_num_features(), self._get_feature_hash_sizes(), self._get_feature_total_num_buckets(),
                                    device if device is not None else self._device, is_sequence=False,
                                    has_feature_processor=any(g.has_feature_processor for g in self._grouped_embedding_configs), need_pos=self._need_pos,
                                    embedding_shard_metadata=self._row_boundaries())

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return self._pooled_lookup(device, fused_params, feature_processor)

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]:
        assert self._pg is not None
        return RwPooledEmbeddingDist(self._pg, self.embedding_dims(), self.qcomm_codecs_registry)

    def create_write_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesWriteDist[KeyedJaggedTensor]:
        assert self._pg is not None
        return RwSparseFeaturesWriteDist(self._pg, self._get_num_features(), self._get_feature_hash_sizes(), device if device is not None else self._device,
                                         embedding_shard_metadata=self._row_boundaries())


class RwSparseFeaturesWriteDist(BaseSparseFeaturesWriteDist[KeyedJaggedTensor]):
    """Embedding UPDATE input dist: (ids, new rows) to the row owners. The rows travel as the KJT's weights, ``D`` floats per id, so ONE KJT all-to-all
    moves both (reference :705-812 sends the values through a second all-to-all)."""

    def __init__(self, pg: dist.ProcessGroup, num_features: int, feature_hash_sizes: List[int], device: Optional[torch.device] = None,
                 embedding_shard_metadata: Optional[List[List[int]]] = None) -> None:
        super().__init__()
        self._pg = pg
        self._world_size = pg.size()
        self._num_features = num_features
        self.register_buffer("_feature_block_size_tensor", torch.tensor([-(-h // self._world_size) for h in feature_hash_sizes], device=device, dtype=torch.int64), persistent=False)
        self._row_pos = [torch.tensor(b, device=device, dtype=torch.int64) for b in embedding_shard_metadata] if embedding_shard_metadata else None
        self._ids_dist = KJTAllToAll(pg=pg, splits=[num_features] * self._world_size)

    def forward(self, sparse_features: KeyedJaggedTensor, embeddings: torch.Tensor) -> Awaitable[Awaitable[KeyedJaggedTensor]]:
        from ..dist_data import TensorAllToAll

        ids, rows = bucketize_embeddings_before_all2all_write(sparse_features, embeddings, self._world_size, self._feature_block_size_tensor, self._row_pos)
        ids_aw = self._ids_dist(ids)
        per_rank = ids.lengths().view(self._world_size, -1).sum(1)
        return _WriteSplitsAwaitable(ids_aw, rows, per_rank, self._pg)


class _WriteSplitsAwaitable(Awaitable[Awaitable[KeyedJaggedTensor]]):
    def __init__(self, ids_aw: Awaitable[Awaitable[KeyedJaggedTensor]], rows: torch.Tensor, per_rank: torch.Tensor, pg: dist.ProcessGroup) -> None:
        super().__init__()
        self._ids_aw, self._rows, self._per_rank, self._pg = ids_aw, rows, per_rank, pg

    def _wait_impl(self) -> Awaitable[KeyedJaggedTensor]:
        return _WriteTensorsAwaitable(self._ids_aw.wait(), self._rows, self._per_rank, self._pg)


class _WriteTensorsAwaitable(Awaitable[KeyedJaggedTensor]):
    def __init__(self, ids_aw: Awaitable[KeyedJaggedTensor], rows: torch.Tensor, per_rank: torch.Tensor, pg: dist.ProcessGroup) -> None:
        super().__init__()
        self._ids_aw, self._rows, self._per_rank, self._pg = ids_aw, rows, per_rank, pg

    def _wait_impl(self) -> KeyedJaggedTensor:
        kjt = self._ids_aw.wait()
        W_ = self._pg.size()
        send = self._per_rank.tolist()
        recv_t = torch.empty(W_, dtype=self._per_rank.dtype, device=self._per_rank.device)
        dist.all_to_all_single(recv_t, self._per_rank.contiguous(), group=self._pg)
        recv = recv_t.tolist()
        D = self._rows.shape[1]
        out = torch.empty(sum(recv), D, dtype=self._rows.dtype, device=self._rows.device)
        dist.all_to_all_single(out, self._rows.contiguous(), output_split_sizes=recv, input_split_sizes=send, group=self._pg)
        # received rows are rank-major (sender, feature, sample); the KJT is key-major after its recat: apply the same permutation to the rows
        F_ = len(kjt.keys())
        lens = kjt.lengths().view(F_, -1)
        B_total = lens.shape[1]
        # source layout of lengths: [sender][feature][sample of sender]; strides per sender
        spr = kjt.stride_per_rank() if hasattr(kjt, "stride_per_rank") and kjt.stride_per_rank() else [B_total // W_] * W_
        src_blocks = []
        col = 0
        for s, b in enumerate(spr):
            src_blocks.append(lens[:, col : col + b])  # [F, b] of sender s
            col += b
        # number of ids per (sender, feature)
        counts = torch.stack([blk.sum(1) for blk in src_blocks])  # [W, F]
        src_off = torch.cumsum(counts.flatten(), 0) - counts.flatten()  # rank-major starts
        order = []
        for f in range(F_):
            for s in range(W_):
                n = int(counts[s, f])
                if n:
                    st = int(src_off[s * F_ + f])
                    order.append(torch.arange(st, st + n, device=out.device))
        rows = out[torch.cat(order)] if order else out
        return KeyedJaggedTensor(keys=kjt.keys(), values=kjt.values(), weights=rows, lengths=kjt.lengths(), stride=kjt.stride())


# ---- inference -----------------------------------------------------------------------------------------------------
class InferRwSparseFeaturesDist(BaseSparseFeaturesDist[InputDistOutputs]):
    """Bucketize on the host device, then copy bucket ``r`` to device ``r``; sequence lookups also get the unbucketize permutation and the bucket
    of every id (reference :834-958)."""

    def __init__(self, world_size: int, num_features: int, feature_hash_sizes: List[int], feature_total_num_buckets: Optional[List[int]] = None,
                 device: Optional[torch.device] = None, is_sequence: bool = False, has_feature_processor: bool = False, need_pos: bool = False,
                 embedding_shard_metadata: Optional[List[List[int]]] = None, keep_original_indices: bool = False) -> None:
        super().__init__()
        self._world_size = world_size
        self._num_features = num_features
        self.feature_block_sizes = [-(-h // world_size) for h in feature_hash_sizes]
        self._tensor_cache: Dict[str, Tuple[torch.Tensor, List[torch.Tensor]]] = {}
        self._embedding_shard_metadata = embedding_shard_metadata
        self._dist = KJTOneToAll([num_features] * world_size, world_size, device)
        self._is_sequence = is_sequence
        self._has_feature_processor = has_feature_processor
        self._need_pos = need_pos
        self._keep_original_indices = keep_original_indices

    def forward(self, sparse_features: KeyedJaggedTensor) -> InputDistOutputs:
        block, row_pos = get_block_sizes_runtime_device(self.feature_block_sizes, sparse_features.device(), self._tensor_cache, self._embedding_shard_metadata, torch.int64)
        bucketized, unbucketize, mapping = bucketize_kjt_inference(
            sparse_features, self._world_size, block, bucketize_pos=(self._has_feature_processor if sparse_features.weights_or_none() is None else self._need_pos),
            block_bucketize_row_pos=row_pos or None, is_sequence=self._is_sequence, keep_original_indices=self._keep_original_indices)
        feats = KJTList(self._dist(bucketized).wait())
        bucketized_lengths = bucketized.lengths().view(self._world_size * self._num_features, -1).sum(1) if self._is_sequence else None
        return InputDistOutputs(features=feats, unbucketize_permute_tensor=unbucketize if self._is_sequence else None,
                                bucket_mapping_tensor=mapping if self._is_sequence else None, bucketized_length=bucketized_lengths)


class InferRwPooledEmbeddingDist(BaseEmbeddingDist[NullShardingContext, List[torch.Tensor], torch.Tensor]):
    """Sum the devices' partial pools on one device (reference :604-641)."""

    def __init__(self, device: torch.device, world_size: int) -> None:
        super().__init__()
        self._dist = EmbeddingsAllToOneReduce(device, world_size)

    def forward(self, local_embs: List[torch.Tensor], sharding_ctx: Optional[NullShardingContext] = None) -> torch.Tensor:
        return self._dist(local_embs)


class InferRwPooledEmbeddingSharding(BaseRwEmbeddingSharding[NullShardingContext, InputDistOutputs, List[torch.Tensor], torch.Tensor]):
    def _copy_weights(self) -> None:
        self._init_rows = {}

    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[InputDistOutputs]:
        return InferRwSparseFeaturesDist(self._world_size, self._get_num_features(), self._get_feature_hash_sizes(), self._get_feature_total_num_buckets(),
                                         device if device is not None else self._device, embedding_shard_metadata=self._row_boundaries())

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return InferGroupedPooledEmbeddingsLookup(self._grouped_embedding_configs_per_rank, self._world_size, fused_params, device, feature_processor,
                                                  device_type_from_sharding_infos=(device.type if device is not None else self._device.type))

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[NullShardingContext, List[torch.Tensor], torch.Tensor]:
        return InferRwPooledEmbeddingDist(device if device is not None else self._device, self._world_size)
