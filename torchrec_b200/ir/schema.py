"""Metadata records the IR serializer stores next to an exported graph, one per module kind it can rebuild.

Counterpart of ``torchrec/ir/schema.py``. Records are plain dataclasses registered by name so ``to_json`` / ``from_json`` round-trip them without
pickling; ``RECORDS`` is what ``ir/serializer.py`` dispatches on.
"""
from __future__ import annotations

import dataclasses
import json
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Type

RECORDS: Dict[str, Type[Any]] = {}


def _record(cls: Type[Any]) -> Type[Any]:
    cls = dataclass(cls)
    RECORDS[cls.__name__] = cls
    return cls


def to_json(record: Any) -> str:
    return json.dumps({"kind": type(record).__name__, "fields": dataclasses.asdict(record)})


def from_json(text: str) -> Any:
    blob = json.loads(text)
    cls = RECORDS[blob["kind"]]
    kw = dict(blob["fields"])
    for f in dataclasses.fields(cls):  # nested table records come back as dicts
        if f.name == "tables":
            kw["tables"] = [EmbeddingBagConfigMetadata(**t) for t in kw["tables"]]
        if f.name == "max_feature_lengths":
            kw["max_feature_lengths"] = [tuple(p) for p in kw["max_feature_lengths"]]
    return cls(**kw)


@_record
class EmbeddingBagConfigMetadata:
    """One table: enough to rebuild its ``EmbeddingBagConfig`` / ``EmbeddingConfig``."""

    name: str
    num_embeddings: int
    embedding_dim: int
    feature_names: List[str]
    data_type: str = "FP32"
    pooling: str = "SUM"
    need_pos: bool = False
    weight_init_min: Optional[float] = None
    weight_init_max: Optional[float] = None


@_record
class EBCMetadata:
    tables: List[EmbeddingBagConfigMetadata]
    is_weighted: bool = False
    device: Optional[str] = None


@_record
class ECMetadata:
    tables: List[EmbeddingBagConfigMetadata]
    need_indices: bool = False
    device: Optional[str] = None


@_record
class FPEBCMetadata:
    features: List[str]
    is_fp_collection: bool = True


@_record
class PositionWeightedModuleMetadata:
    max_feature_length: int


@_record
class PositionWeightedModuleCollectionMetadata:
    max_feature_lengths: List[Tuple[str, int]]


@_record
class KTRegroupAsDictMetadata:
    groups: List[List[str]] = field(default_factory=list)
    keys: List[str] = field(default_factory=list)
