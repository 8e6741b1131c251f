"""Mean squared error, RMSE and R^2.

Reference module: ``torchrec/metrics/mse.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class MSEMetricComputation(_SumStatesComputation):
    STATES = ["error_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        d = predictions.double() - labels.double()
        return {"error_sum": (weights.double() * d * d).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        mse = get("error_sum") / (get("weighted_num_samples") + EPS)
        return [MetricComputationReport(MetricName.MSE, prefix, mse), MetricComputationReport(MetricName.RMSE, prefix, torch.sqrt(mse))]


MSEMetric = _make("MSEMetric", MSEMetricComputation, MetricNamespace.MSE)


def compute_mse(error_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_samples == 0.0, torch.zeros_like(error_sum), error_sum / weighted_num_samples).double()


def compute_rmse(error_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.sqrt(compute_mse(error_sum, weighted_num_samples))


def compute_r_squared(error_sum: torch.Tensor, weighted_num_samples: torch.Tensor, label_sum: torch.Tensor, label_squared_sum: torch.Tensor) -> torch.Tensor:
    total = label_squared_sum - label_sum * label_sum / (weighted_num_samples + EPS)
    return torch.where(total == 0.0, torch.zeros_like(error_sum), 1.0 - error_sum / total).double()


def compute_error_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (weights.double() * (labels.double() - predictions.double()) ** 2).sum(-1)


def get_mse_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    return {"error_sum": compute_error_sum(labels, predictions, weights), "weighted_num_samples": w.sum(-1), "label_sum": (w * y).sum(-1), "label_squared_sum": (w * y * y).sum(-1)}
