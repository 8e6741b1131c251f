"""Unsharded embedding collections (reference torchrec/modules/embedding_modules.py:97-552).

``EmbeddingBagCollection``: KJT -> KeyedTensor of pooled embeddings, one ``nn.EmbeddingBag`` per
table (these are the golden modules the sharded runtime is compared against and the objects
``DistributedModelParallel`` swaps out). ``EmbeddingCollection``: KJT -> Dict[str, JaggedTensor]
of unpooled (sequence) embeddings.
"""
from __future__ import annotations

import abc
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from .embedding_configs import (
    DataType,
    EmbeddingBagConfig,
    EmbeddingConfig,
    PoolingType,
    data_type_to_dtype,
    pooling_type_to_str,
)


def reorder_inverse_indices(inverse_indices: Optional[Tuple[List[str], torch.Tensor]], feature_names: List[str]) -> torch.Tensor:
    if inverse_indices is None:
        return torch.empty(0)
    index_per_name = {name: i for i, name in enumerate(inverse_indices[0])}
    index = torch.tensor([index_per_name[name.split("@")[0]] for name in feature_names], device=inverse_indices[1].device)
    return torch.index_select(inverse_indices[1], 0, index)


def process_pooled_embeddings(pooled_embeddings: List[torch.Tensor], inverse_indices: torch.Tensor) -> torch.Tensor:
    """Concatenate per-feature pooled outputs; for VBE inputs first re-expand each feature to the
    full batch with its inverse indices."""
    if inverse_indices.numel() > 0:
        pooled_embeddings = [emb[inverse_indices[i].long()] for i, emb in enumerate(pooled_embeddings)]
    return torch.cat(pooled_embeddings, dim=1)


def get_embedding_names_by_table(tables: Union[List[EmbeddingBagConfig], List[EmbeddingConfig]]) -> List[List[str]]:
    """Features shared by several tables are disambiguated as ``feature@table``
    (reference embedding_modules.py:75-94)."""
    shared_feature: Dict[str, bool] = {}
    for cfg in tables:
        for name in cfg.feature_names:
            shared_feature[name] = name in shared_feature
    out: List[List[str]] = []
    for cfg in tables:
        out.append([f"{n}@{cfg.name}" if shared_feature[n] else n for n in cfg.feature_names])
    return out


class EmbeddingBagCollectionInterface(abc.ABC, nn.Module):
    @abc.abstractmethod
    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        ...

    @abc.abstractmethod
    def embedding_bag_configs(self) -> List[EmbeddingBagConfig]:
        ...

    @abc.abstractmethod
    def is_weighted(self) -> bool:
        ...


class EmbeddingBagCollection(EmbeddingBagCollectionInterface):
    """Collection of pooled embedding tables.

    ``forward(KJT) -> KeyedTensor`` with values ``[B, sum(dim over (table, feature))]`` and keys in
    table order. ``state_dict`` keys are ``embedding_bags.<table>.weight`` — the layout every
    sharded variant preserves.
    """

    def __init__(self, tables: List[EmbeddingBagConfig], is_weighted: bool = False, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._is_weighted = is_weighted
        self.embedding_bags: nn.ModuleDict = nn.ModuleDict()
        self._embedding_bag_configs = tables
        self._lengths_per_embedding: List[int] = []
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        self._dtypes: List[int] = []
        table_names = set()
        for cfg in tables:
            if cfg.name in table_names:
                raise ValueError(f"Duplicate table name {cfg.name}")
            table_names.add(cfg.name)
            dtype = torch.float32 if cfg.data_type == DataType.FP32 else data_type_to_dtype(cfg.data_type)
            self.embedding_bags[cfg.name] = nn.EmbeddingBag(
                num_embeddings=cfg.num_embeddings, embedding_dim=cfg.embedding_dim, mode=pooling_type_to_str(cfg.pooling),
                device=self._device, include_last_offset=True, dtype=dtype,
            )
            if not cfg.feature_names:
                cfg.feature_names = [cfg.name]
            self._lengths_per_embedding.extend(len(cfg.feature_names) * [cfg.embedding_dim])
        self._embedding_names: List[str] = [n for names in get_embedding_names_by_table(tables) for n in names]
        self._feature_names: List[List[str]] = [t.feature_names for t in tables]
        self.reset_parameters()

    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        flat_feature_names: List[str] = [n for names in self._feature_names for n in names]
        inverse_indices = reorder_inverse_indices(features.inverse_indices_or_none(), flat_feature_names)
        pooled: List[torch.Tensor] = []
        feature_dict = features.to_dict()
        for i, bag in enumerate(self.embedding_bags.values()):
            for name in self._feature_names[i]:
                f = feature_dict[name]
                psw = f.weights() if self._is_weighted else None
                if psw is not None and not torch.is_floating_point(psw):
                    psw = None
                res = bag(input=f.values(), offsets=f.offsets(), per_sample_weights=psw).float()
                pooled.append(res)
        return KeyedTensor(keys=self._embedding_names, values=process_pooled_embeddings(pooled, inverse_indices),
                           length_per_key=self._lengths_per_embedding)

    def is_weighted(self) -> bool:
        return self._is_weighted

    def embedding_bag_configs(self) -> List[EmbeddingBagConfig]:
        return self._embedding_bag_configs

    @property
    def device(self) -> torch.device:
        return self._device

    def reset_parameters(self) -> None:
        if (isinstance(self.device, torch.device) and self.device.type == "meta") or (isinstance(self.device, str) and self.device == "meta"):
            return
        for cfg in self._embedding_bag_configs:
            assert cfg.init_fn is not None
            param = self.embedding_bags[cfg.name].weight
            cfg.init_fn(param)


class EmbeddingCollectionInterface(abc.ABC, nn.Module):
    @abc.abstractmethod
    def forward(self, features: KeyedJaggedTensor) -> Dict[str, JaggedTensor]:
        ...

    @abc.abstractmethod
    def embedding_configs(self) -> List[EmbeddingConfig]:
        ...

    @abc.abstractmethod
    def need_indices(self) -> bool:
        ...

    @abc.abstractmethod
    def embedding_dim(self) -> int:
        ...

    @abc.abstractmethod
    def embedding_names_by_table(self) -> List[List[str]]:
        ...


class EmbeddingCollection(EmbeddingCollectionInterface):
    """Collection of unpooled embedding tables (all tables share one ``embedding_dim``)."""

    def __init__(self, tables: List[EmbeddingConfig], device: Optional[torch.device] = None, need_indices: bool = False, use_gather_select: bool = False) -> None:
        super().__init__()
        # ``use_gather_select``: the sharded collection expands de-duplicated lookups with gather instead of index_select (cheaper backward)
        self._use_gather_select = use_gather_select
        self.embeddings: nn.ModuleDict = nn.ModuleDict()
        self._embedding_configs = tables
        self._embedding_dim: int = -1
        self._need_indices = need_indices
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        table_names = set()
        for cfg in tables:
            if cfg.name in table_names:
                raise ValueError(f"Duplicate table name {cfg.name}")
            table_names.add(cfg.name)
            self._embedding_dim = cfg.embedding_dim if self._embedding_dim < 0 else self._embedding_dim
            if self._embedding_dim != cfg.embedding_dim:
                raise ValueError("All tables in a EmbeddingCollection are required to have same embedding dimension. "
                                 f"Violating case: {cfg.name}'s embedding_dim {cfg.embedding_dim} != {self._embedding_dim}")
            dtype = torch.float32 if cfg.data_type == DataType.FP32 else data_type_to_dtype(cfg.data_type)
            self.embeddings[cfg.name] = nn.Embedding(num_embeddings=cfg.num_embeddings, embedding_dim=cfg.embedding_dim, device=self._device, dtype=dtype)
            if not cfg.feature_names:
                cfg.feature_names = [cfg.name]
        self._embedding_names_by_table: List[List[str]] = get_embedding_names_by_table(tables)
        self._feature_names: List[List[str]] = [t.feature_names for t in tables]
        self.reset_parameters()

    def forward(self, features: KeyedJaggedTensor) -> Dict[str, JaggedTensor]:
        feature_embeddings: Dict[str, JaggedTensor] = {}
        jt_dict = features.to_dict()
        for i, emb_module in enumerate(self.embeddings.values()):
            for j, feature_name in enumerate(self._feature_names[i]):
                embedding_name = self._embedding_names_by_table[i][j]
                f = jt_dict[feature_name]
                lookup = emb_module(input=f.values()).float()
                feature_embeddings[embedding_name] = JaggedTensor(values=lookup, lengths=f.lengths(), weights=f.values() if self._need_indices else None)
        return feature_embeddings

    def need_indices(self) -> bool:
        return self._need_indices

    def embedding_dim(self) -> int:
        return self._embedding_dim

    def embedding_configs(self) -> List[EmbeddingConfig]:
        return self._embedding_configs

    def embedding_names_by_table(self) -> List[List[str]]:
        return self._embedding_names_by_table

    @property
    def device(self) -> torch.device:
        return self._device

    def reset_parameters(self) -> None:
        if self.device.type == "meta":
            return
        for cfg in self._embedding_configs:
            assert cfg.init_fn is not None
            cfg.init_fn(self.embeddings[cfg.name].weight)
