"""Types of the model delta tracker (reference torchrec/distributed/model_tracker/types.py:18-148)."""
from dataclasses import dataclass, field
from enum import Enum
from typing import List, Optional

import torch


@dataclass
class IndexedLookup:
    """ids (global row ids of one table) looked up in batch ``batch_idx`` + optional per-id states."""
    batch_idx: int
    ids: torch.Tensor
    states: Optional[torch.Tensor]
    compact: bool = False


@dataclass
class RawIndexedLookup:
    batch_idx: int
    ids: torch.Tensor
    raw_ids: Optional[torch.Tensor] = None
    runtime_meta: Optional[torch.Tensor] = None


@dataclass
class UniqueRows:
    ids: torch.Tensor
    states: Optional[torch.Tensor]


class TrackingMode(Enum):
    ID_ONLY = "id_only"
    EMBEDDING = "embedding"
    MOMENTUM_LAST = "momentum_last"
    MOMENTUM_DIFF = "momentum_diff"
    ROWWISE_ADAGRAD = "rowwise_adagrad"


class UpdateMode(Enum):
    NONE = "none"
    FIRST = "first"
    LAST = "last"


class Trackers(Enum):
    DELTA_TRACKER = "delta_tracker"
    RAW_ID_TRACKER = "raw_id_tracker"


@dataclass
class RawIdTrackerConfig:
    delete_on_read: bool = True
    fqns_to_skip: List[str] = field(default_factory=list)


@dataclass
class DeltaTrackerConfig:
    tracking_mode: TrackingMode = TrackingMode.ID_ONLY
    consumers: Optional[List[str]] = None
    delete_on_read: bool = True
    auto_compact: bool = False
    fqns_to_skip: List[str] = field(default_factory=list)


@dataclass
class ModelTrackerConfigs:
    raw_id_tracker_config: Optional[RawIdTrackerConfig] = None
    delta_tracker_config: Optional[DeltaTrackerConfig] = None
