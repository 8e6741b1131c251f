"""Weighted mean of labels and predictions.

Reference module: ``torchrec/metrics/average.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class AverageMetricComputation(_SumStatesComputation):
    """Weighted average of the labels (``label_average``) and of the predictions (``prediction_average``)."""

    STATES = ["label_sum", "prediction_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        if predictions is None or weights is None:
            raise RecMetricException("Inputs 'predictions' and 'weights' should not be None for AverageMetricComputation update")
        return get_average_states(labels, predictions, weights)

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.LABEL_AVERAGE, prefix, compute_average(get("label_sum"), get("weighted_num_samples"))),
                MetricComputationReport(MetricName.PREDICTION_AVERAGE, prefix, compute_average(get("prediction_sum"), get("weighted_num_samples")))]


AverageMetric = _make("AverageMetric", AverageMetricComputation, MetricNamespace.AVERAGE)


def compute_average(weighted_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_samples == 0.0, torch.zeros_like(weighted_sum), weighted_sum / weighted_num_samples).double()


def get_average_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    w = weights.double()
    return {"label_sum": (w * labels.double()).sum(-1), "prediction_sum": (w * predictions.double()).sum(-1), "weighted_num_samples": w.sum(-1)}
