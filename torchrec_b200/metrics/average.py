"""Weighted mean of labels and predictions.

Reference module: ``torchrec/metrics/average.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import AverageMetric, AverageMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_average(value_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_samples == 0.0, torch.zeros_like(value_sum), value_sum / weighted_num_samples).double()


def get_average_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    w = weights.double()
    return {"label_sum": (w * labels.double()).sum(-1), "prediction_sum": (w * predictions.double()).sum(-1), "weighted_num_samples": w.sum(-1)}
