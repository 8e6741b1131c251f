"""Structured decision / event logging of the distributed runtime.

Reference: ``torchrec/distributed/logging_handlers.py`` - ``TorchrecComponent`` :37, ``EventLoggingHandler`` :51, ``TrainingOptimizationLogger`` :173 and the
``log_*`` helpers :207-440 (planner results, offloading, ITEP, 2D sharding, kernel changes, pipeline modules, TBE composition, search space).
Events are dict records: kept in memory (``events()``), written as JSON lines when ``TORCHREC_B200_EVENT_LOG`` names a file, and mirrored to the rank-aware
python logger at DEBUG.
"""
from __future__ import annotations

import json
import os
import threading
import time
from enum import Enum
from typing import Any, Dict, List, Optional

from .logger import get_logger
from .logging_utils import EventLoggingHandlerBase, EventScope, EventType, OptimizationTechnique, StackLayer  # noqa: F401

_log = get_logger("torchrec_b200.events")


class TorchrecComponent(Enum):
    PLANNER = "planner"
    ENUMERATOR = "enumerator"
    PROPOSER = "proposer"
    PARTITIONER = "partitioner"
    STORAGE_RESERVATION = "storage_reservation"
    SHARDER = "sharder"
    EMBEDDING_KERNEL = "embedding_kernel"
    TRAIN_PIPELINE = "train_pipeline"
    ITEP = "itep"
    DMP = "dmp"
    CHECKPOINT = "checkpoint"


class EventLoggingHandler(EventLoggingHandlerBase):
    _instance: Optional["EventLoggingHandler"] = None
    _lock = threading.Lock()

    def __init__(self, path: Optional[str] = None, max_events: int = 100_000) -> None:
        self._path = path if path is not None else os.environ.get("TORCHREC_B200_EVENT_LOG")
        self._events: List[Dict[str, Any]] = []
        self._max = max_events
        self._fh = None

    @classmethod
    def get(cls) -> "EventLoggingHandler":
        with cls._lock:
            if cls._instance is None:
                cls._instance = cls()
            return cls._instance

    @classmethod
    def reset(cls, handler: Optional["EventLoggingHandler"] = None) -> None:
        with cls._lock:
            cls._instance = handler

    def log_event(self, component: Any, event_name: str, event_type: EventType = EventType.INFO, technique: OptimizationTechnique = OptimizationTechnique.NONE,
                  scope: EventScope = EventScope.RANK, metadata: Optional[Dict[str, Any]] = None) -> None:
        rec = {"ts": time.time(), "rank": int(os.environ.get("RANK", "0")), "component": getattr(component, "value", str(component)), "event": event_name,
               "type": event_type.value, "technique": technique.value, "scope": scope.value, "metadata": _jsonable(metadata or {})}
        with self._lock:
            if len(self._events) < self._max:
                self._events.append(rec)
            if self._path:
                if self._fh is None:
                    self._fh = open(self._path, "a")
                self._fh.write(json.dumps(rec) + "\n")
        _log.debug("%s/%s %s", rec["component"], event_name, rec["metadata"])

    def events(self, event_name: Optional[str] = None) -> List[Dict[str, Any]]:
        with self._lock:
            return [e for e in self._events if event_name is None or e["event"] == event_name]

    def flush(self) -> None:
        with self._lock:
            if self._fh is not None:
                self._fh.flush()


class TrainingOptimizationLogger(EventLoggingHandler):
    """Handler that keeps only DECISION events of a given technique set - the "why is my job configured like this" trail."""

    def __init__(self, techniques: Optional[List[OptimizationTechnique]] = None, **kw: Any) -> None:
        super().__init__(**kw)
        self._techniques = set(techniques) if techniques else None

    def log_event(self, component: Any, event_name: str, event_type: EventType = EventType.INFO, technique: OptimizationTechnique = OptimizationTechnique.NONE,
                  scope: EventScope = EventScope.RANK, metadata: Optional[Dict[str, Any]] = None) -> None:
        if event_type != EventType.DECISION or (self._techniques is not None and technique not in self._techniques):
            return
        super().log_event(component, event_name, event_type, technique, scope, metadata)


def _jsonable(x: Any) -> Any:
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple, set)):
        return [_jsonable(v) for v in x]
    if isinstance(x, (str, int, float, bool)) or x is None:
        return x
    if isinstance(x, Enum):
        return x.value
    return str(x)


def _emit(component: TorchrecComponent, name: str, technique: OptimizationTechnique = OptimizationTechnique.NONE, event_type: EventType = EventType.DECISION,
          scope: EventScope = EventScope.RANK, **metadata: Any) -> None:
    EventLoggingHandler.get().log_event(component, name, event_type, technique, scope, metadata)


def detect_technique(items: List[Any]) -> OptimizationTechnique:
    """Which memory technique a list of sharding options / plan entries uses (by compute kernel)."""
    kernels = {str(getattr(i, "compute_kernel", "")) for i in items}
    if kernels & {"fused_uvm_caching", "fused_uvm", "key_value", "quant_uvm", "quant_uvm_caching"}:
        return OptimizationTechnique.EMBEDDING_OFFLOADING
    return OptimizationTechnique.NONE


def _plan_rows(best_plan: List[Any]) -> List[Dict[str, Any]]:
    return [{"table": getattr(so, "name", None), "sharding_type": getattr(so, "sharding_type", None), "compute_kernel": getattr(so, "compute_kernel", None),
             "ranks": [getattr(s, "rank", None) for s in getattr(so, "shards", [])], "perf": float(getattr(so, "total_perf", 0.0) or 0.0)} for so in best_plan]


def log_planning_result(best_plan: List[Any], planner_type: str = "", technique: OptimizationTechnique = OptimizationTechnique.NONE, **extra: Any) -> None:
    _emit(TorchrecComponent.PLANNER, "planning_result", technique, planner_type=planner_type, plan=_plan_rows(best_plan), **extra)


def log_offloading_summary(best_plan: List[Any], planner_type: str = "", technique: OptimizationTechnique = OptimizationTechnique.NONE) -> None:
    rows = [r for r in _plan_rows(best_plan) if r["compute_kernel"] in ("fused_uvm", "fused_uvm_caching", "key_value")]
    _emit(TorchrecComponent.PLANNER, "offloading_summary", technique if technique != OptimizationTechnique.NONE else detect_technique(best_plan), planner_type=planner_type,
          offloaded_tables=rows, num_offloaded=len(rows))


def log_storage_reservation(reservation_type: str = "", percentage: Optional[float] = None, dense_storage: Any = None, kjt_storage: Any = None, **extra: Any) -> None:
    _emit(TorchrecComponent.STORAGE_RESERVATION, "storage_reservation", reservation_type=reservation_type, percentage=percentage, dense_storage=dense_storage,
          kjt_storage=kjt_storage, **extra)


def log_planner_config(planner_type: str = "", **config: Any) -> None:
    _emit(TorchrecComponent.PLANNER, "planner_config", planner_type=planner_type, **config)


def log_stats_match(table: str = "", matched: bool = True, **extra: Any) -> None:
    _emit(TorchrecComponent.PLANNER, "stats_match", OptimizationTechnique.EMBEDDING_OFFLOADING, table=table, matched=matched, **extra)


def log_cacheability_resolved(table: str = "", cacheability: Optional[float] = None, **extra: Any) -> None:
    _emit(TorchrecComponent.PROPOSER, "cacheability_resolved", OptimizationTechnique.EMBEDDING_OFFLOADING, table=table, cacheability=cacheability, **extra)


def log_clf_computed(table: str = "", clf: Optional[float] = None, **extra: Any) -> None:
    _emit(TorchrecComponent.PROPOSER, "clf_computed", OptimizationTechnique.EMBEDDING_OFFLOADING, table=table, cache_load_factor=clf, **extra)


def _itep(name: str):
    def fn(**metadata: Any) -> None:
        _emit(TorchrecComponent.ITEP, name, OptimizationTechnique.ITEP, **metadata)

    fn.__name__ = f"log_{name}"
    return fn


log_itep_config = _itep("itep_config")
log_itep_init_state = _itep("itep_init_state")
log_itep_eviction = _itep("itep_eviction")
log_itep_pruning_trigger = _itep("itep_pruning_trigger")
log_itep_checkpoint_save = _itep("itep_checkpoint_save")
log_itep_checkpoint_load = _itep("itep_checkpoint_load")
log_itep_table_config = _itep("itep_table_config")
log_itep_rowwise_shard = _itep("itep_rowwise_shard")
log_itep_ien_pruning_stats = _itep("itep_ien_pruning_stats")
log_itep_ien_pruning_decision = _itep("itep_ien_pruning_decision")
log_itep_itp_info = _itep("itep_itp_info")


def log_two_dim_sharding_config(**metadata: Any) -> None:
    _emit(TorchrecComponent.DMP, "two_dim_sharding_config", OptimizationTechnique.TWO_DIM_SHARDING, **metadata)


def log_kernel_changed(table: str = "", old_kernel: str = "", new_kernel: str = "", reason: str = "", **extra: Any) -> None:
    _emit(TorchrecComponent.EMBEDDING_KERNEL, "kernel_changed", table=table, old_kernel=old_kernel, new_kernel=new_kernel, reason=reason, **extra)


def log_pipeline_module_info(pipeline: str = "", pipelined_modules: Optional[List[str]] = None, **extra: Any) -> None:
    _emit(TorchrecComponent.TRAIN_PIPELINE, "pipeline_module_info", pipeline=pipeline, pipelined_modules=pipelined_modules or [], **extra)


def log_table_assignment(best_plan: List[Any], planner_type: str = "", technique: OptimizationTechnique = OptimizationTechnique.NONE) -> None:
    _emit(TorchrecComponent.PARTITIONER, "table_assignment", technique, planner_type=planner_type, assignment=_plan_rows(best_plan))


def log_table_constraints(constraints: Optional[Dict[str, Any]] = None, planner_type: str = "", technique: OptimizationTechnique = OptimizationTechnique.NONE) -> None:
    _emit(TorchrecComponent.PLANNER, "table_constraints", technique, planner_type=planner_type, constraints={k: str(v) for k, v in (constraints or {}).items()})


def log_tbe_composition(grouped_configs: List[Any], rank: int = 0, technique: OptimizationTechnique = OptimizationTechnique.NONE) -> None:
    groups = [{"tables": g.table_names(), "kernel": getattr(g.compute_kernel, "value", str(g.compute_kernel)), "data_type": str(g.data_type), "dim_sum": g.dim_sum()} for g in grouped_configs]
    _emit(TorchrecComponent.EMBEDDING_KERNEL, "tbe_composition", technique, rank=rank, groups=groups)


def log_search_space_summary(search_space: List[Any], planner_type: str = "", technique: OptimizationTechnique = OptimizationTechnique.NONE) -> None:
    by_type: Dict[str, int] = {}
    for so in search_space:
        by_type[str(getattr(so, "sharding_type", "?"))] = by_type.get(str(getattr(so, "sharding_type", "?")), 0) + 1
    _emit(TorchrecComponent.ENUMERATOR, "search_space_summary", technique, planner_type=planner_type, num_options=len(search_space), by_sharding_type=by_type)


def log_search_space_augmented(*args: Any, **kwargs: Any) -> None:
    _emit(TorchrecComponent.ENUMERATOR, "search_space_augmented", **{f"arg{i}": a for i, a in enumerate(args)}, **kwargs)


def log_proposer_result(proposer: str = "", num_proposals: int = 0, best_perf: Optional[float] = None, **extra: Any) -> None:
    _emit(TorchrecComponent.PROPOSER, "proposer_result", proposer=proposer, num_proposals=num_proposals, best_perf=best_perf, **extra)
