"""Column-wise sharding: a table's columns split into blocks, each block a "table" of its own on some rank.

Reference: ``torchrec/distributed/sharding/cw_sharding.py`` - ``BaseCwEmbeddingSharding`` :61-257 (extends the table-wise base), ``CwPooledEmbeddingSharding``
:260-317, inference variants :320-430. A feature of a column-sharded table is sent to EVERY block owner (the module repeats it in ``feature_names()``);
after the pooled all-to-all the blocks of one feature may be non-adjacent (rank-major order), so the output dist carries a column permutation
(``permute_pooled_embs``) that restores the table's own column order - ``uncombined_*`` describe the blocks, ``embedding_*`` the restored layout.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from ...ops import jagged as J
from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..embedding_sharding import BaseEmbeddingDist, BaseSparseFeaturesDist, C, EmbeddingShardingContext, EmbeddingShardingInfo, F, T, W
from ..embedding_types import BaseEmbeddingLookup, InputDistOutputs, ShardedEmbeddingTable
from ..types import NullShardingContext, ShardMetadata
from .common import make_shard_table, rank_of, shards_of_info
from .tw_sharding import (
    BaseTwEmbeddingSharding,
    InferTwPooledEmbeddingDist,
    InferTwSparseFeaturesDist,
    TwPooledEmbeddingDist,
    TwSparseFeaturesDist,
    _global_md,
)


class BaseCwEmbeddingSharding(BaseTwEmbeddingSharding[C, F, T, W]):
    def __init__(self, sharding_infos: List[EmbeddingShardingInfo], env: Any, device: Optional[torch.device] = None, need_pos: bool = False,
                 qcomm_codecs_registry: Optional[Dict[str, Any]] = None, permute_embeddings: bool = False) -> None:
        super().__init__(sharding_infos, env, device, need_pos, qcomm_codecs_registry)
        self._permute_embeddings = permute_embeddings
        self._permute: Optional[List[int]] = None
        self._combined_names: List[str] = super().embedding_names()
        self._combined_dims: List[int] = super().embedding_dims()
        if permute_embeddings:
            self._init_combined_embeddings()

    def _shard(self, sharding_infos: List[EmbeddingShardingInfo]) -> List[List[ShardedEmbeddingTable]]:
        tables_per_rank: List[List[ShardedEmbeddingTable]] = [[] for _ in range(self._world_size)]
        for info in sharding_infos:
            shards = shards_of_info(info)
            ranks = info.param_sharding.ranks or [rank_of(s.placement) for s in shards]
            gmd = _global_md(info)
            for shard, rank in zip(shards, ranks):
                tables_per_rank[rank].append(make_shard_table(info, shard, shard.shard_sizes[0], shard.shard_sizes[1], gmd))
        return tables_per_rank

    def _init_combined_embeddings(self) -> None:
        """Blocks arrive rank-major. Build (a) the block permutation that groups each embedding name's blocks by ascending column offset, names in order
        of first appearance, and (b) the combined names / dims (reference :176-230)."""
        names = super().embedding_names()
        dims = super().embedding_dims()
        mds = super().embedding_shard_metadata()
        order: Dict[str, List[Tuple[int, int]]] = {}
        for i, (n, md) in enumerate(zip(names, mds)):
            order.setdefault(n, []).append((md.shard_offsets[1] if md is not None else 0, i))
        perm: List[int] = []
        self._combined_names, self._combined_dims = [], []
        for n, blocks in order.items():
            blocks.sort()
            perm.extend(i for _, i in blocks)
            self._combined_names.append(n)
            self._combined_dims.append(sum(dims[i] for _, i in blocks))
        self._permute = None if perm == list(range(len(perm))) else perm
        self._uncombined_dims = dims

    def _permute_callback(self) -> Optional[List[Any]]:
        if not self._permute_embeddings or self._permute is None:
            return None
        dims = self._uncombined_dims
        offsets = [0]
        for d in dims:
            offsets.append(offsets[-1] + d)
        perm = self._permute
        return [lambda t: J.permute_pooled_embs(t, offsets, perm)]

    def embedding_dims(self) -> List[int]:
        return self._combined_dims if self._permute_embeddings else super().embedding_dims()

    def embedding_names(self) -> List[str]:
        return self._combined_names if self._permute_embeddings else super().embedding_names()

    def uncombined_embedding_dims(self) -> List[int]:
        return BaseTwEmbeddingSharding.embedding_dims(self)

    def uncombined_embedding_names(self) -> List[str]:
        return BaseTwEmbeddingSharding.embedding_names(self)


class CwPooledEmbeddingSharding(BaseCwEmbeddingSharding[EmbeddingShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor]):
    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[KeyedJaggedTensor]:
        assert self._pg is not None
        return TwSparseFeaturesDist(self._pg, self.features_per_rank())

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        return self._pooled_lookup(device, fused_params, feature_processor)

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[EmbeddingShardingContext, torch.Tensor, torch.Tensor]:
        assert self._pg is not None
        return TwPooledEmbeddingDist(self._pg, self._dim_sum_per_rank(), self._emb_dim_per_rank_per_feature(), device if device is not None else self._device,
                                     callbacks=self._permute_callback(), qcomm_codecs_registry=self.qcomm_codecs_registry)


class InferCwPooledEmbeddingDist(InferTwPooledEmbeddingDist):
    """reference :369-390."""


class InferCwPooledEmbeddingDistWithPermute(InferTwPooledEmbeddingDist):
    """All-to-one gather followed by the block permutation (reference :399-430)."""

    def __init__(self, device: torch.device, world_size: int, offsets: List[int], permute: List[int]) -> None:
        super().__init__(device, world_size)
        self._offsets, self._perm = offsets, permute

    def forward(self, local_embs: List[torch.Tensor], sharding_ctx: Optional[NullShardingContext] = None) -> torch.Tensor:
        return J.permute_pooled_embs(super().forward(local_embs), self._offsets, self._perm)


class InferCwPooledEmbeddingSharding(BaseCwEmbeddingSharding[NullShardingContext, InputDistOutputs, List[torch.Tensor], torch.Tensor]):
    def _copy_weights(self) -> None:
        self._init_rows = {}

    def create_input_dist(self, device: Optional[torch.device] = None) -> BaseSparseFeaturesDist[InputDistOutputs]:
        return InferTwSparseFeaturesDist(self.features_per_rank(), self._world_size, device)

    def create_lookup(self, device: Optional[torch.device] = None, fused_params: Optional[Dict[str, Any]] = None,
                      feature_processor: Optional[nn.Module] = None) -> BaseEmbeddingLookup:
        from ..embedding_lookup import InferGroupedPooledEmbeddingsLookup

        return InferGroupedPooledEmbeddingsLookup(self._grouped_embedding_configs_per_rank, self._world_size, fused_params, device, feature_processor,
                                                  device_type_from_sharding_infos=(device.type if device is not None else self._device.type))

    def create_output_dist(self, device: Optional[torch.device] = None) -> BaseEmbeddingDist[NullShardingContext, List[torch.Tensor], torch.Tensor]:
        dev = device if device is not None else self._device
        if self._permute_embeddings and self._permute is not None:
            offsets = [0]
            for d in self._uncombined_dims:
                offsets.append(offsets[-1] + d)
            return InferCwPooledEmbeddingDistWithPermute(dev, self._world_size, offsets, self._permute)
        return InferCwPooledEmbeddingDist(dev, self._world_size)
