"""Reference import path ``torchrec/distributed/train_pipeline/postproc.py`` (``NoOpStream`` :23, ``PipelinedPostproc`` :37); implementation in ``utils.py``."""
from __future__ import annotations

from typing import Any

from .utils import PipelinedPostproc  # noqa: F401


class NoOpStream:
    """Stand-in for a CUDA stream on CPU: a context manager whose stream operations do nothing."""

    def __init__(self, stream: Any = None) -> None:
        self._stream = stream

    def __enter__(self) -> "NoOpStream":
        return self

    def __exit__(self, *exc: Any) -> None:
        return None

    def wait_stream(self, other: Any) -> None:
        pass

    def synchronize(self) -> None:
        pass

    def record_event(self, event: Any = None) -> Any:
        return event
