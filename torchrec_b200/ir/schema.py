"""Serializable metadata of the modules the IR serializer knows (reference ``torchrec/ir/schema.py``: ``EmbeddingBagConfigMetadata`` :18, ``EBCMetadata`` :31,
``FPEBCMetadata`` :38, ``PositionWeightedModuleMetadata`` :44, ``PositionWeightedModuleCollectionMetadata`` :49, ``KTRegroupAsDictMetadata`` :54)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple


@dataclass
class EmbeddingBagConfigMetadata:
    num_embeddings: int
    embedding_dim: int
    name: str
    data_type: str
    feature_names: List[str]
    weight_init_max: Optional[float] = None
    weight_init_min: Optional[float] = None
    need_pos: bool = False
    pooling: str = "SUM"


@dataclass
class EBCMetadata:
    tables: List[EmbeddingBagConfigMetadata]
    is_weighted: bool
    device: Optional[str] = None


@dataclass
class ECMetadata:
    tables: List[EmbeddingBagConfigMetadata]
    need_indices: bool = False
    device: Optional[str] = None


@dataclass
class FPEBCMetadata:
    is_fp_collection: bool
    features: List[str]


@dataclass
class PositionWeightedModuleMetadata:
    max_feature_length: int


@dataclass
class PositionWeightedModuleCollectionMetadata:
    max_feature_lengths: List[Tuple[str, int]]


@dataclass
class KTRegroupAsDictMetadata:
    groups: List[List[str]] = field(default_factory=list)
    keys: List[str] = field(default_factory=list)
