"""NE of the serving-time prediction.

Reference module: ``torchrec/metrics/serving_ne.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from typing import Any  # noqa: F401

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .ne import NEMetricComputation  # noqa: F401


class ServingNEMetricComputation(NEMetricComputation):
    """NE over the serving data (examples with weight 0 carry nothing) + the number of examples with a non-zero weight. Reported as
    ``serving_ne-<task>|lifetime_ne`` / ``window_ne`` / ``total_examples``."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._add_state("num_examples", torch.zeros(self._n_tasks, dtype=torch.long), add_window_state=False, dist_reduce_fx="sum", persistent=True)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        if predictions is None or weights is None:
            raise RecMetricException("Inputs 'predictions' and 'weights' should not be None for ServingNEMetricComputation update")
        self.num_examples += torch.count_nonzero(weights, dim=-1).to(self.num_examples.device)
        return super()._batch_states(predictions, labels, weights, **kwargs)

    def _extra_reports(self):
        return [MetricComputationReport(MetricName.TOTAL_EXAMPLES, MetricPrefix.DEFAULT, self.num_examples.detach())]


ServingNEMetric = _make("ServingNEMetric", ServingNEMetricComputation, MetricNamespace.SERVING_NE)
