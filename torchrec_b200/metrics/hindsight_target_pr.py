"""Precision / recall at the threshold that reaches a target precision in hindsight.

Reference module: ``torchrec/metrics/hindsight_target_pr.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import HindsightTargetPRMetric, HindsightTargetPRMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_precision(num_true_positives: torch.Tensor, num_false_positives: torch.Tensor) -> torch.Tensor:
    d = num_true_positives + num_false_positives
    return torch.where(d == 0.0, torch.zeros_like(d), num_true_positives / d).double()


def compute_recall(num_true_positives: torch.Tensor, num_false_negatives: torch.Tensor) -> torch.Tensor:
    d = num_true_positives + num_false_negatives
    return torch.where(d == 0.0, torch.zeros_like(d), num_true_positives / d).double()


def compute_threshold_idx(num_true_positives: torch.Tensor, num_false_positives: torch.Tensor, target_precision: float) -> int:
    """Smallest threshold bucket whose precision reaches the target (last bucket when none does)."""
    ok = (compute_precision(num_true_positives, num_false_positives) >= target_precision).nonzero()
    return int(ok[0]) if ok.numel() else int(num_true_positives.numel() - 1)


def _bucketed(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, thresholds: torch.Tensor, pos_pred: bool, pos_label: bool) -> torch.Tensor:
    pred = (predictions.double().unsqueeze(-1) >= thresholds.double()) == pos_pred
    lab = ((labels.double() >= 0.5) == pos_label).unsqueeze(-1)
    return (weights.double().unsqueeze(-1) * (pred & lab).double()).sum(-2)


def compute_true_pos_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, thresholds: torch.Tensor) -> torch.Tensor:
    return _bucketed(labels, predictions, weights, thresholds, True, True)


def compute_false_pos_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, thresholds: torch.Tensor) -> torch.Tensor:
    return _bucketed(labels, predictions, weights, thresholds, True, False)


def compute_false_neg_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, thresholds: torch.Tensor) -> torch.Tensor:
    return _bucketed(labels, predictions, weights, thresholds, False, True)


def get_pr_states(labels: torch.Tensor, predictions: torch.Tensor, weights: Optional[torch.Tensor], thresholds: torch.Tensor) -> Dict[str, torch.Tensor]:
    if weights is None:
        weights = torch.ones_like(predictions)
    return {"true_pos_sum": compute_true_pos_sum(labels, predictions, weights, thresholds), "false_pos_sum": compute_false_pos_sum(labels, predictions, weights, thresholds),
            "false_neg_sum": compute_false_neg_sum(labels, predictions, weights, thresholds)}
