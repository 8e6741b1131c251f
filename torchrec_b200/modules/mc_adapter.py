"""Turn plain embedding (bag) configs into managed-collision collections in one call (reference torchrec/modules/mc_adapter.py:36-266).

``McEmbeddingCollectionAdapter(tables, input_hash_size, device, world_size, zch_method="mpzch" | "sort_zch")`` builds the tables on the meta
device (the sharder materialises them), one collision module per table (multi-probe ZCH with single-TTL eviction, or sorted ZCH with
distance-LFU eviction) and the wrapping ``ManagedCollisionEmbedding(Bag)Collection``; ``forward`` returns the embeddings and keeps the
remapped ids of the last call in ``remapped_ids``."""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional

import torch
from torch import nn
from torch.nn import Parameter

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .embedding_configs import EmbeddingBagConfig, EmbeddingConfig
from .embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from .hash_mc_modules import HashZchEvictionConfig, HashZchEvictionPolicyName, HashZchManagedCollisionModule
from .mc_embedding_modules import ManagedCollisionEmbeddingBagCollection, ManagedCollisionEmbeddingCollection
from .mc_modules import DistanceLFU_EvictionPolicy, ManagedCollisionCollection, ManagedCollisionModule, MCHManagedCollisionModule


def _collision_modules(configs, input_hash_size: int, device: torch.device, world_size: int, eviction_interval: int, zch_method: str,
                       mpzch_num_buckets: Optional[int], mpzch_max_probe: Optional[int]) -> Dict[str, ManagedCollisionModule]:
    mods: Dict[str, ManagedCollisionModule] = {}
    for cfg in configs:
        if zch_method == "mpzch":
            buckets = mpzch_num_buckets if mpzch_num_buckets else world_size
            if cfg.num_embeddings % buckets != 0:  # uniform buckets need a divisor: fall back to one bucket per rank
                buckets = world_size if cfg.num_embeddings % world_size == 0 else 1
            mods[cfg.name] = HashZchManagedCollisionModule(zch_size=cfg.num_embeddings, device=device, total_num_buckets=buckets, max_probe=int(mpzch_max_probe or 100),
                                                           input_hash_size=input_hash_size, is_inference=False,
                                                           eviction_policy_name=HashZchEvictionPolicyName.SINGLE_TTL_EVICTION,
                                                           eviction_config=HashZchEvictionConfig(features=list(cfg.feature_names), single_ttl=eviction_interval))
        elif zch_method == "sort_zch":
            mods[cfg.name] = MCHManagedCollisionModule(zch_size=cfg.num_embeddings, device=device, input_hash_size=input_hash_size, eviction_interval=max(1, eviction_interval),
                                                       eviction_policy=DistanceLFU_EvictionPolicy())
        else:
            raise NotImplementedError(f"zch method {zch_method!r} is not supported (use 'mpzch' or 'sort_zch')")
    return mods


class McEmbeddingCollectionAdapter(nn.Module):
    def __init__(self, tables: List[EmbeddingConfig], input_hash_size: int, device: torch.device, world_size: int, eviction_interval: int = 1,
                 allow_in_place_embed_weight_update: bool = False, zch_method: str = "", mpzch_num_buckets: Optional[int] = 80, mpzch_max_probe: Optional[int] = 100,
                 embedding_device: Optional[torch.device] = None) -> None:
        super().__init__()
        ec = EmbeddingCollection(tables=tables, device=embedding_device if embedding_device is not None else torch.device("meta"))
        mods = _collision_modules(ec.embedding_configs(), input_hash_size, device, world_size, eviction_interval, zch_method, mpzch_num_buckets, mpzch_max_probe)
        self.mc_embedding_collection = ManagedCollisionEmbeddingCollection(ec, ManagedCollisionCollection(mods, ec.embedding_configs()), return_remapped_features=True)
        self.allow_in_place_embed_weight_update = allow_in_place_embed_weight_update
        self.remapped_ids: Optional[Dict[str, JaggedTensor]] = None

    def forward(self, input: KeyedJaggedTensor) -> Dict[str, JaggedTensor]:
        out, remapped = self.mc_embedding_collection(input)
        self.remapped_ids = remapped.to_dict() if remapped is not None and hasattr(remapped, "to_dict") else remapped
        return out

    def parameters(self, recurse: bool = True) -> Iterator[Parameter]:
        return self.mc_embedding_collection._embedding_collection.parameters(recurse)

    def embedding_bag_configs(self) -> List[EmbeddingConfig]:
        return self.mc_embedding_collection._embedding_collection.embedding_configs()


class McEmbeddingBagCollectionAdapter(nn.Module):
    def __init__(self, tables: List[EmbeddingBagConfig], input_hash_size: int, device: torch.device, world_size: int, eviction_interval: int = 1,
                 allow_in_place_embed_weight_update: bool = False, zch_method: str = "", mpzch_num_buckets: Optional[int] = 80, mpzch_max_probe: Optional[int] = 100,
                 embedding_device: Optional[torch.device] = None) -> None:
        super().__init__()
        ebc = EmbeddingBagCollection(tables=tables, device=embedding_device if embedding_device is not None else torch.device("meta"))
        mods = _collision_modules(ebc.embedding_bag_configs(), input_hash_size, device, world_size, eviction_interval, zch_method, mpzch_num_buckets, mpzch_max_probe)
        self.mc_embedding_bag_collection = ManagedCollisionEmbeddingBagCollection(ebc, ManagedCollisionCollection(mods, ebc.embedding_bag_configs()), return_remapped_features=True)
        self.allow_in_place_embed_weight_update = allow_in_place_embed_weight_update
        self.remapped_ids: Optional[Dict[str, JaggedTensor]] = None

    def forward(self, input_kjt: KeyedJaggedTensor):
        out, remapped = self.mc_embedding_bag_collection(input_kjt)
        self.remapped_ids = remapped.to_dict() if remapped is not None and hasattr(remapped, "to_dict") else remapped
        return out

    def parameters(self, recurse: bool = True) -> Iterator[Parameter]:
        return self.mc_embedding_bag_collection._embedding_bag_collection.parameters(recurse)

    def embedding_bag_configs(self) -> List[EmbeddingBagConfig]:
        return self.mc_embedding_bag_collection._embedding_bag_collection.embedding_bag_configs()
