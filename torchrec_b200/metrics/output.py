"""Statistics of the raw model output.

Reference module: ``torchrec/metrics/output.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from typing import Any  # noqa: F401

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class OutputMetricComputation(RecMetricComputation):
    """Logs model outputs handed over in ``required_inputs``: the batch means of ``latest_imp`` and ``total_latest_imp`` of the latest
    update, reported without a lifetime / window prefix as ``output_latest_imp`` / ``output_total_latest_imp``."""

    REQUIRED = ["latest_imp", "total_latest_imp"]

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        for name in self.REQUIRED:
            self._add_state(name, torch.zeros(self._n_tasks, dtype=torch.double), add_window_state=False, dist_reduce_fx="sum", persistent=False)

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        req = kwargs.get("required_inputs")
        if req is None or not all(k in req for k in self.REQUIRED):
            raise RecMetricException("OutputMetricComputation requires 'latest_imp' and 'total_latest_imp' in kwargs")
        for name in self.REQUIRED:
            v = req[name].float().mean(dim=-1, dtype=torch.double)
            self._buffers[name] = v.reshape(-1).expand(self._n_tasks).clone().to(self._buffers[name].device)

    def _compute(self) -> List[MetricComputationReport]:
        return [MetricComputationReport(MetricName.OUTPUT, MetricPrefix.DEFAULT, self.latest_imp, description="_latest_imp"),
                MetricComputationReport(MetricName.OUTPUT, MetricPrefix.DEFAULT, self.total_latest_imp, description="_total_latest_imp")]


class OutputMetric(RecMetric):
    _namespace: MetricNamespace = MetricNamespace.OUTPUT
    _computation_class = OutputMetricComputation

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._required_inputs.update(OutputMetricComputation.REQUIRED)


