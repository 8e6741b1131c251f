"""Weighted recall at a threshold.

Reference module: ``torchrec/metrics/recall.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import RecallMetric, RecallMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_recall(num_true_positives: torch.Tensor, num_false_negatives: torch.Tensor) -> torch.Tensor:
    d = num_true_positives + num_false_negatives
    return torch.where(d == 0.0, torch.zeros_like(d), num_true_positives / d).double()


def compute_true_pos_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    return (weights.double() * ((predictions.double() >= threshold) & (labels.double() >= 0.5)).double()).sum(-1)


def compute_false_neg_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    return (weights.double() * ((predictions.double() < threshold) & (labels.double() >= 0.5)).double()).sum(-1)


def get_recall_states(labels: torch.Tensor, predictions: torch.Tensor, weights: Optional[torch.Tensor], threshold: float = 0.5) -> Dict[str, torch.Tensor]:
    if weights is None:
        weights = torch.ones_like(predictions)
    return {"true_pos_sum": compute_true_pos_sum(labels, predictions, weights, threshold), "false_neg_sum": compute_false_neg_sum(labels, predictions, weights, threshold)}
