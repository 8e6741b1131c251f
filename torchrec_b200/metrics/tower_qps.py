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
    """Examples per second seen by a tower: sum of the examples of all ranks / the longest elapsed time of any rank, lifetime (after
    ``warmup_steps`` updates) and window, + ``total_examples`` (warm-up included)."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self._warmup_steps = kwargs.pop("warmup_steps", 0)
        super().__init__(*args, **kwargs)
        self._add_state("num_examples", torch.zeros(self._n_tasks, dtype=torch.long), add_window_state=True, dist_reduce_fx="sum", persistent=True)
        self._add_state("warmup_examples", torch.zeros(self._n_tasks, dtype=torch.long), add_window_state=False, dist_reduce_fx="sum", persistent=True)
        self._add_state("time_lapse", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="max", persistent=True)
        self._steps = 0
        self._previous_ts = 0.0

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        self._steps += 1
        count = labels.shape[-1]
        n = torch.full((self._n_tasks,), count, dtype=torch.long, device=self.num_examples.device)
        self.num_examples += n
        ts = time.monotonic()
        if self._steps <= self._warmup_steps:
            self.warmup_examples += n
            if self._steps == self._warmup_steps:
                self._previous_ts = ts
            return
        if self._previous_ts == 0.0:  # no warm-up: the clock starts with the first update, which is not rated
            self.warmup_examples += n
            self._previous_ts = ts
            return
        dt = torch.full((self._n_tasks,), ts - self._previous_ts, dtype=torch.double, device=self.time_lapse.device)
        self.time_lapse += dt
        self._aggregate_window_state("num_examples", n, count)
        self._aggregate_window_state("time_lapse", dt, count)
        self._previous_ts = ts

    def _compute(self) -> List[MetricComputationReport]:
        out = [MetricComputationReport(MetricName.TOWER_QPS, MetricPrefix.LIFETIME, _compute_tower_qps(self.num_examples - self.warmup_examples, self.time_lapse))]
        if self._batch_window_buffers is not None:
            out.append(MetricComputationReport(MetricName.TOWER_QPS, MetricPrefix.WINDOW, _compute_tower_qps(self.get_window_state("num_examples"), self.get_window_state("time_lapse"))))
        out.append(MetricComputationReport(MetricName.TOTAL_EXAMPLES, MetricPrefix.DEFAULT, self.num_examples.detach()))
        return out


def _compute_tower_qps(num_examples: torch.Tensor, time_lapse: torch.Tensor) -> torch.Tensor:
    return torch.where(time_lapse <= 0.0, torch.zeros_like(time_lapse), num_examples.double() / time_lapse.clamp(min=EPS)).double()


TowerQPSMetric = _make("TowerQPSMetric", TowerQPSMetricComputation, MetricNamespace.TOWER_QPS)
