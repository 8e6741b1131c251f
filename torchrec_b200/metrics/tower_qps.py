"""Examples per second of one tower / task.

Reference module: ``torchrec/metrics/tower_qps.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make, _zeros  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class TowerQPSMetricComputation(RecMetricComputation):
    """Examples per second seen by a tower (lifetime and window), max over ranks of the elapsed time."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self._warmup_steps = kwargs.pop("warmup_steps", 0)
        super().__init__(*args, **kwargs)
        self._add_state("num_examples", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="sum", persistent=True)
        self._add_state("time_lapse", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="max", persistent=True)
        self._steps = 0
        self._previous_ts = 0.0

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        self._steps += 1
        if self._steps <= self._warmup_steps:
            return
        ts = time.monotonic()
        if self._steps == self._warmup_steps + 1:
            self._previous_ts = ts
            return
        n = torch.full((self._n_tasks,), float(labels.shape[-1]), dtype=torch.double)
        dt = torch.full((self._n_tasks,), ts - self._previous_ts, dtype=torch.double)
        self.num_examples += n.to(self.num_examples.device)
        self.time_lapse += dt.to(self.time_lapse.device)
        self._aggregate_window_state("num_examples", n, labels.shape[-1])
        self._aggregate_window_state("time_lapse", dt, labels.shape[-1])
        self._previous_ts = ts

    def _compute(self) -> List[MetricComputationReport]:
        out = [MetricComputationReport(MetricName.TOWER_QPS, MetricPrefix.LIFETIME, self.num_examples / (self.time_lapse + EPS))]
        if self._batch_window_buffers is not None:
            out.append(MetricComputationReport(MetricName.TOWER_QPS, MetricPrefix.WINDOW, self.get_window_state("num_examples") / (self.get_window_state("time_lapse") + EPS)))
        return out


TowerQPSMetric = _make("TowerQPSMetric", TowerQPSMetricComputation, MetricNamespace.TOWER_QPS)
