"""Weighted recall at a threshold.

Reference module: ``torchrec/metrics/recall.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class RecallMetricComputation(_SumStatesComputation):
    STATES = ["true_pos_sum", "false_neg_sum"]

    def __init__(self, *args: Any, threshold: float = 0.5, **kwargs: Any) -> None:
        self._threshold = threshold
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        pred = (predictions.double() >= self._threshold).double()
        return {"true_pos_sum": (weights.double() * pred * labels.double()).sum(-1), "false_neg_sum": (weights.double() * (1 - pred) * labels.double()).sum(-1)}

    def _reports(self, get, prefix):
        tp, fn = get("true_pos_sum"), get("false_neg_sum")
        return [MetricComputationReport(MetricName.RECALL, prefix, torch.where(tp + fn == 0.0, torch.zeros_like(tp), tp / (tp + fn)))]


RecallMetric = _make("RecallMetric", RecallMetricComputation, MetricNamespace.RECALL)


def compute_recall(num_true_positives: torch.Tensor, num_false_negitives: torch.Tensor) -> torch.Tensor:  # (sic: the reference's spelling of the keyword)
    d = num_true_positives + num_false_negitives
    return torch.where(d == 0.0, torch.zeros_like(d), num_true_positives / d).double()


def compute_true_pos_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    return (weights.double() * ((predictions.double() >= threshold) & (labels.double() >= 0.5)).double()).sum(-1)


def compute_false_neg_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    return (weights.double() * ((predictions.double() < threshold) & (labels.double() >= 0.5)).double()).sum(-1)


def get_recall_states(labels: torch.Tensor, predictions: torch.Tensor, weights: Optional[torch.Tensor], threshold: float = 0.5) -> Dict[str, torch.Tensor]:
    if weights is None:
        weights = torch.ones_like(predictions)
    return {"true_pos_sum": compute_true_pos_sum(labels, predictions, weights, threshold), "false_neg_sum": compute_false_neg_sum(labels, predictions, weights, threshold)}
