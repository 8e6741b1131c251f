"""Normalized entropy.

Reference module: ``torchrec/metrics/ne.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import NEMetric, NEMetricComputation, compute_ne  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_cross_entropy(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> torch.Tensor:
    p = torch.clamp(predictions.double(), eta, 1 - eta)
    return -weights.double() * (labels.double() * torch.log2(p) + (1.0 - labels.double()) * torch.log2(1.0 - p))


def compute_logloss(ce_sum: torch.Tensor, pos_labels: torch.Tensor, neg_labels: torch.Tensor, eta: float) -> torch.Tensor:
    """Mean natural-log loss from the base-2 cross entropy sum."""
    n = pos_labels + neg_labels
    n = torch.where(n == 0.0, torch.full_like(n, eta), n)
    return torch.log(torch.tensor(2.0, dtype=ce_sum.dtype, device=ce_sum.device)) * ce_sum / n


def get_ne_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    return {"cross_entropy_sum": compute_cross_entropy(labels, predictions, weights, eta).sum(-1), "weighted_num_samples": w.sum(-1),
            "pos_labels": (w * y).sum(-1), "neg_labels": (w * (1.0 - y)).sum(-1)}
