"""Sharders of the single-process fused collections (reference torchrec/distributed/fused_embeddingbag.py:41-170,
fused_embedding.py). A FusedEmbeddingBagCollection already names its optimizer, so sharding it is sharding the plain
collection with that optimizer turned into ``fused_params``."""
from typing import Any, Dict, Iterator, List, Optional, Type

import torch
from torch import nn

from ..modules.fused_embedding_modules import FusedEmbeddingBagCollection, FusedEmbeddingCollection, _optim_type
from .embedding import EmbeddingCollectionSharder, ShardedEmbeddingCollection
from .embedding_types import BaseEmbeddingSharder
from .embeddingbag import EmbeddingBagCollectionSharder, ShardedEmbeddingBagCollection
from .types import ParameterSharding, ShardingEnv


def _fused_params_of(module, base: Optional[Dict[str, Any]]) -> Dict[str, Any]:
    kw = dict(module._optimizer_kwargs)
    fp = dict(base or {})
    fp["optimizer"] = _optim_type(module._optimizer_type)
    if "lr" in kw:
        fp["learning_rate"] = kw["lr"]
    for k in ("eps", "weight_decay", "momentum"):
        if k in kw:
            fp[k] = kw[k]
    if "betas" in kw:
        fp["beta1"], fp["beta2"] = kw["betas"]
    return fp


class ShardedFusedEmbeddingBagCollection(ShardedEmbeddingBagCollection):
    def __init__(self, module: FusedEmbeddingBagCollection, table_name_to_parameter_sharding: Dict[str, ParameterSharding], env: ShardingEnv,
                 fused_params: Optional[Dict[str, Any]] = None, device: Optional[torch.device] = None, qcomm_codecs_registry=None, module_fqn: Optional[str] = None) -> None:
        super().__init__(module, table_name_to_parameter_sharding, env, _fused_params_of(module, fused_params), device, qcomm_codecs_registry, module_fqn)


class FusedEmbeddingBagCollectionSharder(BaseEmbeddingSharder[FusedEmbeddingBagCollection]):
    def shard(self, module: FusedEmbeddingBagCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedFusedEmbeddingBagCollection:
        return ShardedFusedEmbeddingBagCollection(module, params, env, self.fused_params, device, self.qcomm_codecs_registry, module_fqn)

    def shardable_parameters(self, module: FusedEmbeddingBagCollection) -> Dict[str, nn.Parameter]:
        return {name.split(".")[-2]: p for name, p in module.named_parameters() if name.endswith(".weight")}

    @property
    def module_type(self) -> Type[FusedEmbeddingBagCollection]:
        return FusedEmbeddingBagCollection

    def sharding_types(self, compute_device_type: str) -> List[str]:
        from .types import ShardingType

        return [t for t in super().sharding_types(compute_device_type) if t != ShardingType.DATA_PARALLEL.value]


# ---- moved to ``fused_embedding.py`` (their reference import path); still importable from here ----
_MOVED_TO_FUSED_EMBEDDING = ('ShardedFusedEmbeddingCollection', 'FusedEmbeddingCollectionSharder')


def __getattr__(name: str):
    if name in _MOVED_TO_FUSED_EMBEDDING:
        from . import fused_embedding as _m

        return getattr(_m, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
