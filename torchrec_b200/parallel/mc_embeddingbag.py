"""Sharded ManagedCollisionEmbeddingBagCollection (reference torchrec/distributed/mc_embeddingbag.py:46-223)."""
from typing import Dict, List, Optional, Type

import torch
from torch import nn

from ..modules.mc_embedding_modules import ManagedCollisionEmbeddingBagCollection
from .embedding_types import BaseEmbeddingSharder
from .embeddingbag import EmbeddingBagCollectionSharder
from .mc_embedding_modules import BaseShardedManagedCollisionEmbeddingCollection
from .mc_modules import ManagedCollisionCollectionSharder
from .types import ParameterSharding, ShardingEnv, ShardingType


class ShardedManagedCollisionEmbeddingBagCollection(BaseShardedManagedCollisionEmbeddingCollection):
    @property
    def _embedding_bag_collection(self):
        return self._embedding_module


class ManagedCollisionEmbeddingBagCollectionSharder(BaseEmbeddingSharder[ManagedCollisionEmbeddingBagCollection]):
    def __init__(self, ebc_sharder: Optional[EmbeddingBagCollectionSharder] = None, mc_sharder: Optional[ManagedCollisionCollectionSharder] = None,
                 fused_params=None, qcomm_codecs_registry=None) -> None:
        super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._e_sharder = ebc_sharder or EmbeddingBagCollectionSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._mc_sharder = mc_sharder or ManagedCollisionCollectionSharder()

    def shard(self, module: ManagedCollisionEmbeddingBagCollection, params: Dict[str, ParameterSharding], env: ShardingEnv,
              device: Optional[torch.device] = None, module_fqn: Optional[str] = None) -> ShardedManagedCollisionEmbeddingBagCollection:
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedManagedCollisionEmbeddingBagCollection(module, params, self._e_sharder, self._mc_sharder, env, device, module_fqn)

    def shardable_parameters(self, module: ManagedCollisionEmbeddingBagCollection) -> Dict[str, nn.Parameter]:
        return self._e_sharder.shardable_parameters(module._embedding_module)

    @property
    def module_type(self) -> Type[ManagedCollisionEmbeddingBagCollection]:
        return ManagedCollisionEmbeddingBagCollection

    def sharding_types(self, compute_device_type: str) -> List[str]:
        # the remap service decouples the ZCH partition from the table layout: every model-parallel type works
        return [t for t in self._e_sharder.sharding_types(compute_device_type) if t != ShardingType.DATA_PARALLEL.value]
