"""Sum of the sample weights.

Reference module: ``torchrec/metrics/sum_weights.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SingleSumComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class SumWeightsMetricComputation(_SingleSumComputation):
    """sum w. Parity: sum_weights.py:21-95."""

    STATES = ["weighted_sum"]
    NAME = MetricName.SUM_WEIGHTS

    def _sum(self, predictions, labels, weights):
        return weights.double().sum(-1)


SumWeightsMetric = _make("SumWeightsMetric", SumWeightsMetricComputation, MetricNamespace.SUM_WEIGHTS)


def compute_weighted_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return weights.double().sum(-1)


def get_weighted_sum_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"weighted_sum": compute_weighted_sum(labels, predictions, weights)}
