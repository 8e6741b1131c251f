"""Normalized entropy over positive samples.

Reference module: ``torchrec/metrics/ne_positive.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import NEPositiveMetric, NEPositiveMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_cross_entropy_positive(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> torch.Tensor:
    return -weights.double() * labels.double() * torch.log2(torch.clamp(predictions.double(), eta, 1 - eta))


def compute_ne_positive(ce_positive_sum: torch.Tensor, weighted_num_samples: torch.Tensor, pos_labels: torch.Tensor, neg_labels: torch.Tensor, eta: float) -> torch.Tensor:
    mean_label = pos_labels / (weighted_num_samples + EPS)
    return ce_positive_sum / (-pos_labels * torch.log2(mean_label + eta) + EPS)


def get_ne_positive_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    return {"cross_entropy_positive_sum": compute_cross_entropy_positive(labels, predictions, weights, eta).sum(-1), "weighted_num_samples": w.sum(-1),
            "pos_labels": (w * y).sum(-1), "neg_labels": (w * (1.0 - y)).sum(-1)}
