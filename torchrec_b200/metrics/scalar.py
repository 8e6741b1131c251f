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
    """Logs a scalar that travels in the labels slot (a loss term, a learning rate, ...): lifetime = the value of the latest batch,
    window = the mean of the per-batch values inside the window; averaged over the ranks."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._add_state("labels", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="mean", persistent=False)
        self._add_state("window_count", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="mean", persistent=False)

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        n = labels.shape[-1]
        value = labels.double().mean(-1).to(self.labels.device)
        one = torch.ones(self._n_tasks, dtype=torch.double, device=self.labels.device)
        self._buffers["labels"] = value
        self._buffers["window_count"] = one
        self._aggregate_window_state("labels", value, n)
        self._aggregate_window_state("window_count", one, n)

    def _compute(self) -> List[MetricComputationReport]:
        out = [MetricComputationReport(MetricName.SCALAR, MetricPrefix.LIFETIME, self.labels)]
        if self._batch_window_buffers is not None:
            out.append(MetricComputationReport(MetricName.SCALAR, MetricPrefix.WINDOW, self.get_window_state("labels") / self.get_window_state("window_count").clamp(min=1.0)))
        return out


ScalarMetric = _make("ScalarMetric", ScalarMetricComputation, MetricNamespace.SCALAR)
