"""Work items of the asynchronous metric pipelines (reference metrics/metric_job_types.py:18-95)."""
from __future__ import annotations

import concurrent.futures
from typing import Any, Dict

import torch

from .metric_state_snapshot import MetricStateSnapshot


class MetricUpdateJob:
    """One ``update(model_out)`` call handed to the CPU worker (D2H transfer happens there or was started by the producer)."""

    __slots__ = ["model_out", "kwargs", "merged_count"]

    def __init__(self, model_out: Dict[str, torch.Tensor], kwargs: Dict[str, Any], merged_count: int = 1) -> None:
        self.model_out, self.kwargs, self.merged_count = model_out, kwargs, merged_count


class MetricComputeJob:
    """Cross-rank sync + compute over a state snapshot; the result is published through ``future``."""

    __slots__ = ["future", "metric_state_snapshot"]

    def __init__(self, future: "concurrent.futures.Future[Dict[str, Any]]", metric_state_snapshot: MetricStateSnapshot) -> None:
        self.future, self.metric_state_snapshot = future, metric_state_snapshot


class SynchronizationMarker:
    """Placed in the update queue where a compute was requested: everything enqueued before it is part of that compute."""

    __slots__ = ["future"]

    def __init__(self, future: "concurrent.futures.Future[Dict[str, Any]]") -> None:
        self.future = future
