"""Mean squared error, RMSE and R^2.

Reference module: ``torchrec/metrics/mse.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import MSEMetric, MSEMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

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
