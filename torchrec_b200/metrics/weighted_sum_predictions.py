"""Weighted sum of the predictions.

Reference module: ``torchrec/metrics/weighted_sum_predictions.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SingleSumComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class WeightedSumPredictionsMetricComputation(_SingleSumComputation):
    """sum w * prediction (NaN predictions count 0). Parity: weighted_sum_predictions.py:21-99."""

    STATES = ["weighted_predictions_sum"]
    NAME = MetricName.WEIGHTED_SUM_PREDICTIONS

    def _needs(self):
        return ["predictions"]

    def _sum(self, predictions, labels, weights):
        return (weights.double() * torch.nan_to_num(predictions.double(), 0.0)).sum(-1)


WeightedSumPredictionsMetric = _make("WeightedSumPredictionsMetric", WeightedSumPredictionsMetricComputation, MetricNamespace.WEIGHTED_SUM_PREDICTIONS)


def compute_weighted_predictions_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (weights.double() * torch.nan_to_num(predictions.double(), 0.0)).sum(-1)


def get_weighted_sum_prediction_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"weighted_sum_predictions": compute_weighted_predictions_sum(labels, predictions, weights)}
