"""Scalar metrics of the hash ZCH tables (hit / insert / collision / eviction rates) and their sinks.

Reference: ``torchrec/modules/hash_mc_metrics.py`` - ``ScalarLoggerBackend`` :21, ``ConsoleScalarLoggerBackend`` :53, ``ScalarLogger`` :80. ``ScalarLogger`` itself
(the accumulating nn.Module) lives with the ZCH module in ``hash_mc_modules.py``."""
from __future__ import annotations

import abc
import logging
from typing import Dict, List, Optional

from .hash_mc_modules import ScalarLogger  # noqa: F401

logger = logging.getLogger(__name__)


class ScalarLoggerBackend(abc.ABC):
    """Where scalars go (console, TensorBoard-like service, a test list)."""

    @abc.abstractmethod
    def log(self, run_type: str, step: int, scalars: Dict[str, float]) -> None:
        ...

    def flush(self) -> None:
        pass


class ConsoleScalarLoggerBackend(ScalarLoggerBackend):
    def __init__(self, every_n_steps: int = 1, log_file_path: str = "") -> None:
        self._every = max(1, every_n_steps)
        if log_file_path:  # the records also go to a file
            handler = logging.FileHandler(log_file_path, mode="w")
            handler.setFormatter(logging.Formatter("%(asctime)s %(message)s"))
            logger.addHandler(handler)
            if logger.level == logging.NOTSET or logger.level > logging.INFO:
                logger.setLevel(logging.INFO)

    def log(self, run_type: str, step: int, scalars: Dict[str, float]) -> None:
        if step % self._every == 0:
            logger.info("[%s] step %d: %s", run_type, step, ", ".join(f"{k}={v:.6g}" for k, v in sorted(scalars.items())))


class ListScalarLoggerBackend(ScalarLoggerBackend):
    """Keeps the records in memory (tests, notebooks)."""

    def __init__(self) -> None:
        self.records: List[Dict[str, object]] = []

    def log(self, run_type: str, step: int, scalars: Dict[str, float]) -> None:
        self.records.append({"run_type": run_type, "step": step, **scalars})
