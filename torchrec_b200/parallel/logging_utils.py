"""Event-logging vocabulary (reference ``torchrec/distributed/logging_utils.py``: ``EventType`` :17, ``StackLayer`` :27, ``OptimizationTechnique`` :37,
``EventScope`` :49, ``EventLoggingHandlerBase`` :58)."""
from __future__ import annotations

import abc
from enum import Enum
from typing import Any, Dict, Optional


class EventType(Enum):
    INFO = "info"
    WARNING = "warning"
    ERROR = "error"
    METRIC = "metric"
    DECISION = "decision"


class StackLayer(Enum):
    PLANNER = "planner"
    SHARDING = "sharding"
    KERNEL = "kernel"
    PIPELINE = "pipeline"
    MODULE = "module"
    CHECKPOINT = "checkpoint"


class OptimizationTechnique(Enum):
    NONE = "none"
    UVM_OFFLOADING = "uvm_offloading"
    EMBEDDING_OFFLOADING = "embedding_offloading"
    ITEP = "itep"
    TWO_DIM_SHARDING = "two_dim_sharding"
    FULLY_SHARDED = "fully_sharded"
    QUANTIZED_COMMS = "quantized_comms"
    PREFETCH_PIPELINE = "prefetch_pipeline"


class EventScope(Enum):
    JOB = "job"
    RANK = "rank"
    MODULE = "module"
    TABLE = "table"


class EventLoggingHandlerBase(abc.ABC):
    """Sink of structured events; subclasses decide where they go (python logging, a file, a metrics service)."""

    @abc.abstractmethod
    def log_event(self, component: Any, event_name: str, event_type: EventType = EventType.INFO, technique: OptimizationTechnique = OptimizationTechnique.NONE,
                  scope: EventScope = EventScope.RANK, metadata: Optional[Dict[str, Any]] = None) -> None:
        ...

    def flush(self) -> None:
        pass
