"""Sharded ManagedCollisionEmbeddingCollection (reference torchrec/distributed/mc_embedding.py:41-139)."""
from typing import Dict, List, Optional, Type

import torch
from torch import nn

from ..modules.mc_embedding_modules import ManagedCollisionEmbeddingCollection
from .embedding import EmbeddingCollectionSharder
from .embedding_types import BaseEmbeddingSharder
from .mc_embedding_modules import BaseShardedManagedCollisionEmbeddingCollection
from .mc_modules import ManagedCollisionCollectionSharder
from .types import ParameterSharding, ShardingEnv, ShardingType


class ShardedManagedCollisionEmbeddingCollection(BaseShardedManagedCollisionEmbeddingCollection):
    @property
    def _embedding_collection(self):
        return self._embedding_module


class ManagedCollisionEmbeddingCollectionSharder(BaseEmbeddingSharder[ManagedCollisionEmbeddingCollection]):
    def __init__(self, ec_sharder: Optional[EmbeddingCollectionSharder] = None, mc_sharder: Optional[ManagedCollisionCollectionSharder] = None,
                 fused_params=None, qcomm_codecs_registry=None) -> None:
        super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._e_sharder = ec_sharder or EmbeddingCollectionSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._mc_sharder = mc_sharder or ManagedCollisionCollectionSharder()

    def shard(self, module: ManagedCollisionEmbeddingCollection, params: Dict[str, ParameterSharding], env: ShardingEnv,
              device: Optional[torch.device] = None, module_fqn: Optional[str] = None) -> ShardedManagedCollisionEmbeddingCollection:
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedManagedCollisionEmbeddingCollection(module, params, self._e_sharder, self._mc_sharder, env, device, module_fqn)

    def shardable_parameters(self, module: ManagedCollisionEmbeddingCollection) -> Dict[str, nn.Parameter]:
        return self._e_sharder.shardable_parameters(module._embedding_module)

    @property
    def module_type(self) -> Type[ManagedCollisionEmbeddingCollection]:
        return ManagedCollisionEmbeddingCollection

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [t for t in self._e_sharder.sharding_types(compute_device_type) if t != ShardingType.DATA_PARALLEL.value]
