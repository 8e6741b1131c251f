"""Hook point for logging ITEP pruning events (reference torchrec/modules/pruning_logger.py:19-49): a context manager around each
pruning pass; the default does nothing, deployments subclass it to publish to their telemetry."""
from __future__ import annotations

from abc import ABC, abstractmethod
from contextlib import contextmanager
from types import SimpleNamespace
from typing import Generator, Optional


class PruningLogBase(object):
    pass


class PruningLogger(ABC):
    @classmethod
    @abstractmethod
    @contextmanager
    def pruning_logger(cls, event: str, trainer: Optional[str] = None, publisher: Optional[str] = None) -> Generator[object, None, None]:
        ...


class PruningLoggerDefault(PruningLogger):
    """No-op logger: yields a scratch namespace the caller may fill with statistics."""

    @classmethod
    @contextmanager
    def pruning_logger(cls, event: str, trainer: Optional[str] = None, publisher: Optional[str] = None) -> Generator[object, None, None]:
        yield SimpleNamespace()
