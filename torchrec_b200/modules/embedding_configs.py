"""Embedding table configuration types (reference torchrec/modules/embedding_configs.py:348-457)."""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum, unique
from functools import partial
from math import sqrt
from typing import Callable, Dict, List, NamedTuple, Optional

import torch

from ..types import DataType


@unique
class PoolingType(Enum):
    SUM = "SUM"
    MEAN = "MEAN"
    NONE = "NONE"


# MEAN pooling under row-wise style shardings is computed as SUM + a divisor applied after the
# reduce-scatter (reference embedding_configs.py:97-115)
def pooling_type_to_pooling_mode(pooling_type: PoolingType, sharding_type: Optional[str] = None):
    from ..ops.tbe import PoolingMode

    if pooling_type == PoolingType.SUM:
        return PoolingMode.SUM
    if pooling_type == PoolingType.MEAN:
        if sharding_type in ("row_wise", "table_row_wise", "grid_shard"):
            return PoolingMode.SUM
        return PoolingMode.MEAN
    if pooling_type == PoolingType.NONE:
        return PoolingMode.NONE
    raise ValueError(f"unsupported pooling type {pooling_type}")


def pooling_type_to_str(pooling_type: PoolingType) -> str:
    return {PoolingType.SUM: "sum", PoolingType.MEAN: "mean"}[pooling_type]


DATA_TYPE_NUM_BITS: Dict[DataType, int] = {
    DataType.FP32: 32, DataType.FP16: 16, DataType.BF16: 16, DataType.INT8: 8, DataType.UINT8: 8,
    DataType.INT4: 4, DataType.INT2: 2, DataType.FP8: 8, DataType.INT32: 32, DataType.INT64: 64,
}


def dtype_to_data_type(dtype: torch.dtype) -> DataType:
    m = {
        torch.float32: DataType.FP32, torch.float: DataType.FP32, torch.float16: DataType.FP16, torch.half: DataType.FP16,
        torch.bfloat16: DataType.BF16, torch.int64: DataType.INT64, torch.long: DataType.INT64, torch.int32: DataType.INT32,
        torch.int: DataType.INT32, torch.int8: DataType.INT8, torch.uint8: DataType.UINT8, torch.quint8: DataType.UINT8,
        torch.qint8: DataType.INT8, torch.quint4x2: DataType.INT4, torch.quint2x4: DataType.INT2,
    }
    if dtype in m:
        return m[dtype]
    raise Exception(f"Invalid data type {dtype}")


def data_type_to_dtype(data_type: DataType) -> torch.dtype:
    m = {
        DataType.FP32: torch.float32, DataType.FP16: torch.float16, DataType.BF16: torch.bfloat16, DataType.INT64: torch.int64,
        DataType.INT32: torch.int32, DataType.INT8: torch.int8, DataType.UINT8: torch.uint8, DataType.INT4: torch.quint4x2,
        DataType.INT2: torch.quint2x4, DataType.FP8: torch.float8_e4m3fn,
    }
    if data_type in m:
        return m[data_type]
    raise ValueError(f"DataType {data_type} cannot be converted to dtype")


# ---- virtual (key-value backed) tables: eviction policies (reference embedding_configs.py:170-345) -----------------------------------
@dataclass
class VirtualTableEvictionPolicy:
    """Base of the eviction policies of a virtual table: a table whose id space is larger than its resident rows and whose rows live
    in a key-value store with an HBM cache in front (compute kernel ``key_value``). ``meta_header_len`` = elements of the per-row
    header (8 B key, 4 B timestamp, 1 bit used + 31 bit count) in units of the table's element size."""

    meta_header_len: int = 0
    embedding_dim: int = 0
    initialized: bool = False

    def init_metaheader_config(self, data_type: DataType, embedding_dim: int) -> None:
        if self.initialized:
            return
        self.meta_header_len = 16 // data_type_to_dtype(data_type).itemsize
        self.embedding_dim = embedding_dim
        self.initialized = True

    def get_meta_header_len(self) -> int:
        return self.meta_header_len

    def get_embedding_dim(self) -> int:
        return self.embedding_dim


@dataclass
class CountBasedEvictionPolicy(VirtualTableEvictionPolicy):
    training_id_eviction_trigger_count: int = 0  # ids per rank that trigger an eviction pass
    eviction_threshold: int = 15                 # rows seen fewer times than this are evicted (0: never)
    decay_rate: float = 0.99
    inference_eviction_threshold: Optional[int] = None

    def __post_init__(self) -> None:
        if self.inference_eviction_threshold is None:
            self.inference_eviction_threshold = self.eviction_threshold


@dataclass
class FeatureScoreBasedEvictionPolicy(VirtualTableEvictionPolicy):
    decay_rate: float = 0.99
    training_id_eviction_trigger_count: int = 0
    training_id_keep_count: int = 0
    eviction_ttl_mins: int = 0
    max_inference_id_num_per_rank: int = 0
    inference_eviction_feature_score_threshold: Optional[float] = None
    feature_score_mapping: Optional[Dict[str, float]] = None  # feature -> score weight
    feature_score_default_value: Optional[float] = None
    enable_auto_feature_score_collection: bool = False


@dataclass
class FeatureScoreMapping:
    """Per-feature score weights of a virtual table (feature-score based eviction) and whether eviction is on."""

    feature_score_mapping: Dict[str, float] = field(default_factory=dict)
    eviction_enabled: bool = False


@dataclass
class TimestampBasedEvictionPolicy(VirtualTableEvictionPolicy):
    training_id_eviction_trigger_count: int = 0
    eviction_ttl_mins: int = 24 * 60
    inference_eviction_ttl_mins: Optional[int] = None

    def __post_init__(self) -> None:
        if self.inference_eviction_ttl_mins is None:
            self.inference_eviction_ttl_mins = self.eviction_ttl_mins


@dataclass
class CountTimestampMixedEvictionPolicy(VirtualTableEvictionPolicy):
    training_id_eviction_trigger_count: int = 0
    eviction_threshold: int = 15
    decay_rate: float = 0.99
    eviction_ttl_mins: int = 24 * 60
    inference_eviction_threshold: Optional[int] = None
    inference_eviction_ttl_mins: Optional[int] = None

    def __post_init__(self) -> None:
        if self.inference_eviction_threshold is None:
            self.inference_eviction_threshold = self.eviction_threshold
        if self.inference_eviction_ttl_mins is None:
            self.inference_eviction_ttl_mins = self.eviction_ttl_mins


@dataclass
class FeatureL2NormBasedEvictionPolicy(VirtualTableEvictionPolicy):
    training_id_eviction_trigger_count: int = 0
    eviction_threshold: float = 0.0
    inference_eviction_threshold: Optional[float] = None

    def __post_init__(self) -> None:
        if self.inference_eviction_threshold is None:
            self.inference_eviction_threshold = self.eviction_threshold


@dataclass
class NoEvictionPolicy(VirtualTableEvictionPolicy):
    pass


def eviction_policy_to_cache_algorithm(policy: Optional[VirtualTableEvictionPolicy]) -> str:
    """Which victim selection of the HBM row cache (``ops/uvm.py``) realises a policy: count / score based -> LFU, time based -> LRU."""
    if isinstance(policy, (CountBasedEvictionPolicy, FeatureScoreBasedEvictionPolicy, FeatureL2NormBasedEvictionPolicy)):
        return "lfu"
    return "lru"


@dataclass
class BaseEmbeddingConfig:
    num_embeddings: int
    embedding_dim: int
    name: str = ""
    data_type: DataType = DataType.FP32
    feature_names: List[str] = field(default_factory=list)
    weight_init_max: Optional[float] = None
    weight_init_min: Optional[float] = None
    num_embeddings_post_pruning: Optional[int] = None
    init_fn: Optional[Callable[[torch.Tensor], Optional[torch.Tensor]]] = None
    # when the position-weighted feature processor wraps the table
    need_pos: bool = False
    input_dim: Optional[int] = None
    total_num_buckets: Optional[int] = None
    use_virtual_table: bool = False
    virtual_table_eviction_policy: Optional[VirtualTableEvictionPolicy] = None
    enable_embedding_update: bool = False
    stash_weights: bool = False

    def get_weight_init_max(self) -> float:
        if self.weight_init_max is None:
            return sqrt(1 / self.num_embeddings)
        return self.weight_init_max

    def get_weight_init_min(self) -> float:
        if self.weight_init_min is None:
            return -sqrt(1 / self.num_embeddings)
        return self.weight_init_min

    def num_features(self) -> int:
        return len(self.feature_names)

    def __post_init__(self) -> None:
        if self.init_fn is None:
            self.init_fn = partial(torch.nn.init.uniform_, a=self.get_weight_init_min(), b=self.get_weight_init_max())
        if self.use_virtual_table and self.virtual_table_eviction_policy is not None:
            self.virtual_table_eviction_policy.init_metaheader_config(self.data_type, self.embedding_dim)


@dataclass
class EmbeddingTableConfig(BaseEmbeddingConfig):
    pooling: PoolingType = PoolingType.SUM
    is_weighted: bool = False
    has_feature_processor: bool = False
    embedding_names: List[str] = field(default_factory=list)


@dataclass
class EmbeddingBagConfig(BaseEmbeddingConfig):
    """Config of a pooled table (``nn.EmbeddingBag`` semantics)."""

    pooling: PoolingType = PoolingType.SUM


@dataclass
class EmbeddingConfig(BaseEmbeddingConfig):
    """Config of an unpooled (sequence) table (``nn.Embedding`` semantics)."""


class QuantConfig(NamedTuple):
    activation: object
    weight: object
    per_table_weight_dtype: Optional[Dict[str, torch.dtype]] = None


def data_type_to_sparse_type(data_type: DataType) -> str:
    """Row format name of a table data type. The reference returns FBGEMM's ``SparseType`` enum; this framework has no FBGEMM and its
    kernels key on the lower-case format name (``"fp32"``, ``"fp16"``, ``"bf16"``, ``"int8"``, ``"int4"``, ``"int2"``, ``"fp8"``)."""
    names = {"FP32": "fp32", "FP16": "fp16", "BF16": "bf16", "INT8": "int8", "UINT8": "int8", "INT4": "int4", "INT2": "int2", "FP8": "fp8", "NFP8": "fp8"}
    key = data_type.name if hasattr(data_type, "name") else str(data_type)
    if key not in names:
        raise ValueError(f"Invalid DataType {data_type}")
    return names[key]
