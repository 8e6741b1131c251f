"""Metric module that does nothing (reference metrics/noop_metric_module.py:20-76): lets trainer code call the metric API
unconditionally when metrics are disabled (e.g. on non-logging ranks or in throughput benchmarks)."""
from __future__ import annotations

from concurrent.futures import Future
from typing import Any, Dict, List, Optional

import torch

from .deferrable_metrics import DeferrableMetrics
from .metric_module import RecMetricModule


class NoOpMetricModule(RecMetricModule):
    def __init__(self) -> None:
        torch.nn.Module.__init__(self)
        self.trained_batches = 0

    def _update_rec_metrics(self, model_out: Dict[str, torch.Tensor], **kwargs: Any) -> None:
        return None

    def update(self, model_out: Dict[str, torch.Tensor], **kwargs: Any) -> None:
        return None

    def should_compute(self) -> bool:
        return False

    def compute(self) -> DeferrableMetrics:  # type: ignore[override]
        return DeferrableMetrics()

    def local_compute(self) -> DeferrableMetrics:  # type: ignore[override]
        return DeferrableMetrics()

    def sync(self) -> None:
        return None

    def unsync(self) -> None:
        return None

    def reset(self) -> None:
        return None

    def get_required_inputs(self) -> Optional[List[str]]:
        return None

    def get_pre_compute_states(self, pg: Any = None) -> Dict[str, Any]:
        return {}

    def load_pre_compute_states(self, source: Dict[str, Any]) -> None:
        return None

    def shutdown(self) -> None:
        return None

    def async_compute(self, future: "Optional[Future[Dict[str, Any]]]" = None) -> "Future[Dict[str, Any]]":
        future = future or Future()
        future.set_result({})
        return future
