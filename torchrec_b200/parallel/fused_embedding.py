"""Reference import path ``torchrec/distributed/fused_embedding.py`` (``ShardedFusedEmbeddingCollection`` :31, ``FusedEmbeddingCollectionSharder`` :93);
both live next to their bag twins in ``fused_embeddingbag.py``."""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Type
import torch
from torch import nn
from ..modules.fused_embedding_modules import FusedEmbeddingBagCollection, FusedEmbeddingCollection, _optim_type
from .embedding import EmbeddingCollectionSharder, ShardedEmbeddingCollection
from .embedding_types import BaseEmbeddingSharder
from .embeddingbag import EmbeddingBagCollectionSharder, ShardedEmbeddingBagCollection
from .types import ParameterSharding, ShardingEnv
from .fused_embeddingbag import _fused_params_of  # noqa: F401


class ShardedFusedEmbeddingCollection(ShardedEmbeddingCollection):
    def __init__(self, module: FusedEmbeddingCollection, table_name_to_parameter_sharding: Dict[str, ParameterSharding], env: ShardingEnv,
                 fused_params: Optional[Dict[str, Any]] = None, device: Optional[torch.device] = None, qcomm_codecs_registry=None, **kw: Any) -> None:
        super().__init__(module, table_name_to_parameter_sharding, env, _fused_params_of(module, fused_params), device, qcomm_codecs_registry, **kw)


class FusedEmbeddingCollectionSharder(BaseEmbeddingSharder[FusedEmbeddingCollection]):
    def shard(self, module: FusedEmbeddingCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedFusedEmbeddingCollection:
        return ShardedFusedEmbeddingCollection(module, params, env, self.fused_params, device, self.qcomm_codecs_registry)

    def shardable_parameters(self, module: FusedEmbeddingCollection) -> Dict[str, nn.Parameter]:
        return {name: h.weight for name, h in module.embeddings.items()}

    @property
    def module_type(self) -> Type[FusedEmbeddingCollection]:
        return FusedEmbeddingCollection

    def sharding_types(self, compute_device_type: str) -> List[str]:
        from .types import ShardingType

        return [t for t in super().sharding_types(compute_device_type) if t != ShardingType.DATA_PARALLEL.value]
