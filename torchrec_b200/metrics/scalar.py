"""A scalar the trainer logs through the metric module (window / lifetime averaged).

Reference module: ``torchrec/metrics/scalar.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make, _zeros  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class ScalarMetricComputation(RecMetricComputation):
    """Reports the last observed scalar (labels carry the value) and its window average."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._add_state("labels", _zeros(self._n_tasks), add_window_state=False, dist_reduce_fx="max", persistent=False)
        self._add_state("window_count", _zeros(self._n_tasks), add_window_state=False, dist_reduce_fx="sum", persistent=False)
        self._add_state("window_sum", _zeros(self._n_tasks), add_window_state=False, dist_reduce_fx="sum", persistent=False)

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        self.labels = labels.double().mean(-1).to(self.labels.device)
        self.window_count += 1
        self.window_sum += labels.double().mean(-1).to(self.window_sum.device)

    def _compute(self) -> List[MetricComputationReport]:
        return [MetricComputationReport(MetricName.SCALAR, MetricPrefix.LIFETIME, self.labels),
                MetricComputationReport(MetricName.SCALAR, MetricPrefix.WINDOW, self.window_sum / (self.window_count + EPS))]


ScalarMetric = _make("ScalarMetric", ScalarMetricComputation, MetricNamespace.SCALAR)
