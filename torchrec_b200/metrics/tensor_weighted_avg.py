"""Weighted average of a named tensor of the model output.

Reference module: ``torchrec/metrics/tensor_weighted_avg.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class TensorWeightedAvgMetricComputation(_SumStatesComputation):
    """Weighted average of an arbitrary named tensor from ``required_inputs``."""

    STATES = ["weighted_sum", "weighted_num_samples"]

    def __init__(self, *args: Any, tensor_name: Optional[str] = None, weighted: bool = True, description: Optional[str] = None, **kwargs: Any) -> None:
        self._tensor_name = tensor_name
        self._weighted = weighted
        self._description = description
        super().__init__(*args, **kwargs)

    def _needs(self):
        return []

    def _batch_states(self, predictions, labels, weights, **kwargs):
        t = kwargs.get("required_inputs", {}).get(self._tensor_name) if self._tensor_name else predictions
        if t is None:
            raise RecMetricException(f"TensorWeightedAvg needs required input '{self._tensor_name}'")
        t = t.reshape(1, -1).double()
        w = weights.double() if self._weighted else torch.ones_like(t)
        return {"weighted_sum": (t * w).sum(-1), "weighted_num_samples": w.sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.TENSOR_WEIGHTED_AVG, prefix, get("weighted_sum") / (get("weighted_num_samples") + EPS), description=self._description)]


TensorWeightedAvgMetric = _make("TensorWeightedAvgMetric", TensorWeightedAvgMetricComputation, MetricNamespace.TENSOR_WEIGHTED_AVG)


def get_mean(value_sum: torch.Tensor, num_samples: torch.Tensor) -> torch.Tensor:
    return value_sum / (num_samples + EPS)
