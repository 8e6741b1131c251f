"""Recall@k of multi-class predictions.

Reference module: ``torchrec/metrics/multiclass_recall.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import MulticlassRecallMetric, MulticlassRecallMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_true_positives_at_k(predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, n_classes: int) -> torch.Tensor:
    """``[n_classes]``: entry k-1 = weighted number of samples whose label is among the top-k classes."""
    ranks = torch.argsort(predictions, dim=-1, descending=True)
    hit_at = (ranks == labels.long().unsqueeze(-1)).double()  # [N, C] one-hot of the label's rank
    return (torch.cumsum(hit_at, dim=-1) * weights.double().unsqueeze(-1)).sum(0)


def compute_multiclass_recall_at_k(tp_at_k: torch.Tensor, total_weights: torch.Tensor) -> torch.Tensor:
    return tp_at_k / (total_weights.unsqueeze(-1) + EPS) if tp_at_k.dim() > total_weights.dim() else tp_at_k / (total_weights + EPS)


def get_multiclass_recall_states(predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, n_classes: int) -> Dict[str, torch.Tensor]:
    return {"tp_at_k": compute_true_positives_at_k(predictions, labels, weights, n_classes), "total_weights": weights.double().sum(-1)}
