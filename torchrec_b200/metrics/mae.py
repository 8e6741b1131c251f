"""Mean absolute error.

Reference module: ``torchrec/metrics/mae.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class MAEMetricComputation(_SumStatesComputation):
    STATES = ["error_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"error_sum": (weights.double() * (predictions.double() - labels.double()).abs()).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.MAE, prefix, get("error_sum") / (get("weighted_num_samples") + EPS))]


MAEMetric = _make("MAEMetric", MAEMetricComputation, MetricNamespace.MAE)


def compute_mae(error_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_samples == 0.0, torch.zeros_like(error_sum), error_sum / weighted_num_samples).double()


def compute_error_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (weights.double() * (labels.double() - predictions.double()).abs()).sum(-1)


def get_mae_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"error_sum": compute_error_sum(labels, predictions, weights), "weighted_num_samples": weights.double().sum(-1)}
