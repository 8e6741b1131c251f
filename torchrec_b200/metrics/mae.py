"""Mean absolute error.

Reference module: ``torchrec/metrics/mae.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import MAEMetric, MAEMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_mae(error_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_samples == 0.0, torch.zeros_like(error_sum), error_sum / weighted_num_samples).double()


def compute_error_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (weights.double() * (labels.double() - predictions.double()).abs()).sum(-1)


def get_mae_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"error_sum": compute_error_sum(labels, predictions, weights), "weighted_num_samples": weights.double().sum(-1)}
