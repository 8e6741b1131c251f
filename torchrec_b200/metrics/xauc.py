"""Cross AUC for regression: weighted fraction of pairs ordered like their labels.

Reference module: ``torchrec/metrics/xauc.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import XAUCMetric, XAUCMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_error_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """Weighted count of pairs (i < j) whose prediction order agrees with their label order."""
    out = []
    for t in range(predictions.shape[0]):
        p, y, w = predictions[t].double(), labels[t].double(), weights[t].double()
        agree = (torch.sign(p[:, None] - p[None, :]) == torch.sign(y[:, None] - y[None, :])).double()
        out.append(torch.triu(agree * (w[:, None] * w[None, :]), diagonal=1).sum())
    return torch.stack(out)


def compute_weighted_num_pairs(weights: torch.Tensor) -> torch.Tensor:
    w = weights.double()
    return torch.stack([torch.triu(w[t][:, None] * w[t][None, :], diagonal=1).sum() for t in range(w.shape[0])])


def compute_xauc(error_sum: torch.Tensor, weighted_num_pairs: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_pairs == 0.0, torch.zeros_like(error_sum), error_sum / weighted_num_pairs).double()


def get_xauc_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"error_sum": compute_error_sum(labels, predictions, weights), "weighted_num_pairs": compute_weighted_num_pairs(weights)}
