"""Grid sharding: column blocks (like CW), each block row-split inside one host (like TWRW).

Reference: ``torchrec/distributed/sharding/grid_sharding.py`` - ``BaseGridEmbeddingSharding`` :67-344, ``GridPooledEmbeddingDist`` :347-473,
``GridPooledEmbeddingSharding`` :476-560. Mechanics = TWRW's two-hop output dist (host reduce-scatter, cross-host all-to-all) over "tables" that are
column blocks, plus CW's column permutation that puts a feature's blocks back in the table's own column order.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch

from ...ops import jagged as J
from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..embedding_sharding import C, EmbeddingShardingContext, EmbeddingShardingInfo, F, T, W
from ..types import QuantizedCommCodecs, ShardingEnv
from .twrw_sharding import BaseTwRwEmbeddingSharding, TwRwPooledEmbeddingDist, TwRwPooledEmbeddingSharding, TwRwSparseFeaturesDist


class BaseGridEmbeddingSharding(BaseTwRwEmbeddingSharding[C, F, T, W]):
    """``_shard`` of the TWRW base already cuts by column block first; this class adds the combined (per table) view of names and dims."""

    def __init__(self, sharding_infos: List[EmbeddingShardingInfo], env: ShardingEnv, device: Optional[torch.device] = None, need_pos: bool = False,
                 qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__(sharding_infos, env, device, need_pos, qcomm_codecs_registry)
        self._init_combined_embeddings()

    def _init_combined_embeddings(self) -> None:
        names = BaseTwRwEmbeddingSharding.embedding_names(self)
        dims = BaseTwRwEmbeddingSharding.embedding_dims(self)
        mds = BaseTwRwEmbeddingSharding.embedding_shard_metadata(self)
        order: Dict[str, List[Tuple[int, int]]] = {}
        for i, (n, md) in enumerate(zip(names, mds)):
            order.setdefault(n, []).append((md.shard_offsets[1] if md is not None else 0, i))
        perm: List[int] = []
        self._combined_names: List[str] = []
        self._combined_dims: List[int] = []
        for n, blocks in order.items():
            blocks.sort()
            perm.extend(i for _, i in blocks)
            self._combined_names.append(n)
            self._combined_dims.append(sum(dims[i] for _, i in blocks))
        self._permute: Optional[List[int]] = None if perm == list(range(len(perm))) else perm
        self._uncombined_dims = dims

    def embedding_dims(self) -> List[int]:
        return self._combined_dims

    def embedding_names(self) -> List[str]:
        return self._combined_names

    def uncombined_embedding_dims(self) -> List[int]:
        return BaseTwRwEmbeddingSharding.embedding_dims(self)

    def uncombined_embedding_names(self) -> List[str]:
        return BaseTwRwEmbeddingSharding.embedding_names(self)


GridSparseFeaturesDist = TwRwSparseFeaturesDist


class GridPooledEmbeddingDist(TwRwPooledEmbeddingDist):
    """reference :347-473 (same two hops as TWRW)."""


class GridPooledEmbeddingSharding(BaseGridEmbeddingSharding[EmbeddingShardingContext, KeyedJaggedTensor, torch.Tensor, torch.Tensor], TwRwPooledEmbeddingSharding):
    def _output_callbacks(self) -> Optional[List[Any]]:
        if self._permute is None:
            return None
        offsets = [0]
        for d in self._uncombined_dims:
            offsets.append(offsets[-1] + d)
        perm = self._permute
        return [lambda t: J.permute_pooled_embs(t, offsets, perm)]
