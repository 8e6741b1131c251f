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
