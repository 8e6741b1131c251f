"""Mean squared error, RMSE and R^2.

Reference module: ``torchrec/metrics/mse.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from typing import Any  # noqa: F401

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class MSEMetricComputation(_SumStatesComputation):
    """Weighted MSE and RMSE; ``include_r_squared`` adds the coefficient of determination."""

    STATES = ["error_sum", "weighted_num_samples", "label_sum", "label_squared_sum"]

    def __init__(self, *args: Any, include_r_squared: bool = False, **kwargs: Any) -> None:
        self._include_r_squared = include_r_squared
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        if predictions is None or weights is None:
            raise RecMetricException("Inputs 'predictions' and 'weights' should not be None for MSEMetricComputation update")
        return get_mse_states(labels, predictions, weights)

    def _reports(self, get, prefix):
        out = [MetricComputationReport(MetricName.MSE, prefix, compute_mse(get("error_sum"), get("weighted_num_samples"))),
               MetricComputationReport(MetricName.RMSE, prefix, compute_rmse(get("error_sum"), get("weighted_num_samples")))]
        if self._include_r_squared:
            out.append(MetricComputationReport(MetricName.R_SQUARED, prefix, compute_r_squared(get("error_sum"), get("weighted_num_samples"), get("label_sum"), get("label_squared_sum"))))
        return out


MSEMetric = _make("MSEMetric", MSEMetricComputation, MetricNamespace.MSE)


def compute_mse(error_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_samples == 0.0, torch.zeros_like(error_sum), error_sum / weighted_num_samples).double()


def compute_rmse(error_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.sqrt(compute_mse(error_sum, weighted_num_samples))


def compute_r_squared(error_sum: torch.Tensor, weighted_num_samples: torch.Tensor, label_sum: torch.Tensor, label_squared_sum: torch.Tensor) -> torch.Tensor:
    """1 - SS_res / SS_tot with SS_tot = sum w y^2 - (sum w y)^2 / sum w; 1 when the labels do not vary."""
    n = torch.where(weighted_num_samples == 0.0, torch.ones_like(weighted_num_samples), weighted_num_samples).double()
    total = label_squared_sum.double() - label_sum.double() * label_sum.double() / n
    safe = torch.where(total == 0.0, torch.ones_like(total), total)
    return torch.where(total == 0.0, torch.ones_like(total), 1.0 - error_sum.double() / safe).double()


def compute_error_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (weights.double() * (labels.double() - predictions.double()) ** 2).sum(-1)


def get_mse_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    return {"error_sum": compute_error_sum(labels, predictions, weights), "weighted_num_samples": w.sum(-1), "label_sum": (w * y).sum(-1), "label_squared_sum": (w * y * y).sum(-1)}
