"""Rank-aware logging helpers (reference torchrec/distributed/logger.py, logging_handlers.py, utils EventLoggingHandler)."""
from __future__ import annotations

import functools
import logging
import os
import time
from typing import Any, Callable, Dict, Optional, TypeVar

import torch.distributed as dist

_F = TypeVar("_F", bound=Callable[..., Any])
_log_handlers: Dict[str, logging.Handler] = {"default": logging.NullHandler()}


def _rank() -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank()
    return int(os.environ.get("RANK", "0"))


class LazyStr:
    """Defers building an expensive log string until a handler actually formats it."""

    def __init__(self, fn: Callable[[], str]) -> None:
        self._fn = fn

    def __str__(self) -> str:
        return self._fn()

    __repr__ = __str__


def get_logger(name: str = "torchrec_b200", all_ranks: bool = False) -> logging.Logger:
    """rank-0-only logger by default; ``all_ranks`` prefixes records with the rank instead of dropping them."""
    logger = logging.getLogger(name + (".all" if all_ranks else ".rank0"))
    if not getattr(logger, "_trb_configured", False):
        class _F(logging.Filter):
            def filter(self, record: logging.LogRecord) -> bool:
                if all_ranks:
                    record.msg = f"[rank {_rank()}] {record.msg}"
                    return True
                return _rank() == 0

        logger.addFilter(_F())
        logger._trb_configured = True  # type: ignore[attr-defined]
    return logger


class CappedLogger:
    """Logs the first ``cap`` occurrences of every message key, then goes quiet (hot-path warnings)."""

    def __init__(self, logger: Optional[logging.Logger] = None, cap: int = 5) -> None:
        self._logger = logger or get_logger()
        self._cap = cap
        self._counts: Dict[str, int] = {}

    def warning(self, key: str, msg: Any = None) -> None:
        n = self._counts.get(key, 0)
        self._counts[key] = n + 1
        if n < self._cap:
            self._logger.warning(msg if msg is not None else key)
        elif n == self._cap:
            self._logger.warning(f"(suppressing further '{key}' messages)")


def _torchrec_method_logger(**wrapper_kwargs: Any) -> Callable[[_F], _F]:
    """Decorator: log entry / exceptions / wall time of a method at DEBUG level on the framework logger."""

    def decorator(func: _F) -> _F:
        @functools.wraps(func)
        def wrapper(*args: Any, **kwargs: Any) -> Any:
            log = logging.getLogger("torchrec_b200.method")
            t0 = time.perf_counter()
            try:
                result = func(*args, **kwargs)
            except BaseException as e:
                log.error(LazyStr(lambda: f"{func.__qualname__} raised {type(e).__name__}: {e}"))
                raise
            log.debug(LazyStr(lambda: f"{func.__qualname__} ok in {(time.perf_counter() - t0) * 1e3:.2f} ms"))
            return result

        return wrapper  # type: ignore[return-value]

    return decorator


class EventLoggingHandler:
    """Pluggable sink of structured framework events (plan built, module sharded, checkpoint, ...)."""

    _sinks: Dict[str, Callable[[str, Dict[str, Any]], None]] = {}

    @classmethod
    def register(cls, name: str, sink: Callable[[str, Dict[str, Any]], None]) -> None:
        cls._sinks[name] = sink

    @classmethod
    def emit(cls, event: str, **fields: Any) -> None:
        for sink in cls._sinks.values():
            sink(event, fields)

    @classmethod
    def event_logger(cls, event: str) -> Callable[[_F], _F]:
        def decorator(func: _F) -> _F:
            @functools.wraps(func)
            def wrapper(*args: Any, **kwargs: Any) -> Any:
                t0 = time.perf_counter()
                ok = True
                try:
                    return func(*args, **kwargs)
                except BaseException:
                    ok = False
                    raise
                finally:
                    cls.emit(event, function=func.__qualname__, success=ok, duration_ms=(time.perf_counter() - t0) * 1e3)

            return wrapper  # type: ignore[return-value]

        return decorator


# ---- percentile logger, planner decision logs, multiprocess debugger (reference logging_handlers.py:207-441, utils.py:578-597) ----
class PercentileLogger:
    """Collects samples (e.g. per-step latencies, wait counters) and logs p50/p90/p99 every ``log_every`` additions."""

    def __init__(self, name: str, log_every: int = 1000, max_samples: int = 100000, logger: "Optional[logging.Logger]" = None) -> None:
        import logging as _logging

        self.name, self.log_every, self.max_samples = name, log_every, max_samples
        self._samples: "List[float]" = []
        self._n = 0
        self._logger = logger or _logging.getLogger(__name__)

    def add(self, value: float) -> None:
        self._n += 1
        if len(self._samples) < self.max_samples:
            self._samples.append(float(value))
        else:
            self._samples[self._n % self.max_samples] = float(value)
        if self._n % self.log_every == 0:
            self._logger.info("%s", self.summary())

    def percentiles(self, qs=(50, 90, 99)) -> "Dict[int, float]":
        if not self._samples:
            return {q: float("nan") for q in qs}
        s = sorted(self._samples)
        return {q: s[min(len(s) - 1, int(round(q / 100.0 * (len(s) - 1))))] for q in qs}

    def summary(self) -> str:
        p = self.percentiles()
        return f"[{self.name}] n={self._n} " + " ".join(f"p{q}={v:.4g}" for q, v in p.items())


def log_planner_config(planner: object, logger: "Optional[logging.Logger]" = None) -> "Dict[str, object]":
    """Structured record of how a planner was configured (topology, proposers, partitioner, constraints)."""
    import logging as _logging

    topo = getattr(planner, "_topology", None)
    rec = {
        "planner": type(planner).__name__,
        "world_size": getattr(topo, "world_size", None), "local_world_size": getattr(topo, "local_world_size", None),
        "compute_device": getattr(topo, "compute_device", None),
        "hbm_cap": getattr(getattr(topo, "devices", [None])[0], "storage", None) and topo.devices[0].storage.hbm,
        "proposers": [type(p).__name__ for p in getattr(planner, "_proposers", [])],
        "partitioner": type(getattr(planner, "_partitioner", None)).__name__,
        "constraints": sorted((getattr(planner, "_constraints", None) or {}).keys()),
    }
    (logger or _logging.getLogger(__name__)).info("planner config: %s", rec)
    return rec


def log_table_assignment(plan: object, logger: "Optional[logging.Logger]" = None) -> "List[Dict[str, object]]":
    """One record per table of a ShardingPlan: sharding type, compute kernel, ranks."""
    import logging as _logging

    out = []
    for path, mod_plan in getattr(plan, "plan", {}).items():
        for table, ps in mod_plan.items():
            out.append({"module": path, "table": table, "sharding_type": getattr(ps, "sharding_type", None),
                        "compute_kernel": getattr(ps, "compute_kernel", None), "ranks": list(getattr(ps, "ranks", None) or [])})
    lg = logger or _logging.getLogger(__name__)
    for r in out:
        lg.info("table assignment: %s", r)
    return out


class ForkedPdb:
    """``pdb`` usable inside a forked / spawned rank: re-opens /dev/stdin for the session (reference distributed/utils.py:578-597).
    ``ForkedPdb().set_trace()`` in a rank, attach from the launching terminal."""

    def __new__(cls, *args, **kwargs):
        import pdb
        import sys

        class _Forked(pdb.Pdb):
            def interaction(self, *a, **k):
                stdin = sys.stdin
                try:
                    sys.stdin = open("/dev/stdin")
                    pdb.Pdb.interaction(self, *a, **k)
                finally:
                    sys.stdin = stdin

        return _Forked(*args, **kwargs)
