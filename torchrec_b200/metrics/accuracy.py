"""Weighted accuracy at a threshold.

Reference module: ``torchrec/metrics/accuracy.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import AccuracyMetric, AccuracyMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_accuracy(accuracy_sum: torch.Tensor, weighted_num_samples: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_samples == 0.0, torch.zeros_like(accuracy_sum), accuracy_sum / weighted_num_samples).double()


def compute_accuracy_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    hit = ((predictions.double() >= threshold) == (labels.double() >= 0.5)).double()
    return (weights.double() * hit).sum(-1)


def get_accuracy_states(labels: torch.Tensor, predictions: torch.Tensor, weights: Optional[torch.Tensor], threshold: float = 0.5) -> Dict[str, torch.Tensor]:
    if weights is None:
        weights = torch.ones_like(predictions)
    return {"accuracy_sum": compute_accuracy_sum(labels, predictions, weights, threshold), "weighted_num_samples": weights.double().sum(-1)}
