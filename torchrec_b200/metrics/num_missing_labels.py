"""Count of samples whose label is missing (negative sentinel).

Reference module: ``torchrec/metrics/num_missing_labels.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SingleSumComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class NumMissingLabelsMetricComputation(_SingleSumComputation):
    """sum of weights of samples whose label is NaN. Parity: num_missing_labels.py:21-95."""

    STATES = ["missing_label_sum"]
    NAME = MetricName.NUM_MISSING_LABELS

    def _sum(self, predictions, labels, weights):
        return torch.where(torch.isnan(labels), weights.double(), torch.zeros_like(weights, dtype=torch.double)).sum(-1)


NumMissingLabelsMetric = _make("NumMissingLabelsMetric", NumMissingLabelsMetricComputation, MetricNamespace.NUM_MISSING_LABELS)


def compute_missing_label_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (labels.double() < 0).double().sum(-1)


def get_num_missing_labels_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"missing_label_sum": compute_missing_label_sum(labels, predictions, weights)}
