"""Weighted accuracy at a threshold.

Reference module: ``torchrec/metrics/accuracy.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class AccuracyMetricComputation(_SumStatesComputation):
    STATES = ["accuracy_sum", "weighted_num_samples"]

    def __init__(self, *args: Any, threshold: float = 0.5, **kwargs: Any) -> None:
        self._threshold = threshold
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        pred = (predictions.double() >= self._threshold).double()
        return {"accuracy_sum": (weights.double() * (pred == labels.double()).double()).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.ACCURACY, prefix, get("accuracy_sum") / (get("weighted_num_samples") + EPS))]


AccuracyMetric = _make("AccuracyMetric", AccuracyMetricComputation, MetricNamespace.ACCURACY)


def compute_accuracy(accuracy_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_samples == 0.0, torch.zeros_like(accuracy_sum), accuracy_sum / weighted_num_samples).double()


def compute_accuracy_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    hit = ((predictions.double() >= threshold) == (labels.double() >= 0.5)).double()
    return (weights.double() * hit).sum(-1)


def get_accuracy_states(labels: torch.Tensor, predictions: torch.Tensor, weights: Optional[torch.Tensor], threshold: float = 0.5) -> Dict[str, torch.Tensor]:
    if weights is None:
        weights = torch.ones_like(predictions)
    return {"accuracy_sum": compute_accuracy_sum(labels, predictions, weights, threshold), "weighted_num_samples": weights.double().sum(-1)}
