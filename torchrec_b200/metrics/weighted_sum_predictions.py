"""Weighted sum of the predictions.

Reference module: ``torchrec/metrics/weighted_sum_predictions.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import WeightedSumPredictionsMetric, WeightedSumPredictionsMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_weighted_predictions_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (weights.double() * torch.nan_to_num(predictions.double(), 0.0)).sum(-1)


def get_weighted_sum_prediction_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"weighted_sum_predictions": compute_weighted_predictions_sum(labels, predictions, weights)}
