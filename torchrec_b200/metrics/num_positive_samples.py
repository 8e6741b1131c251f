"""Weighted number of positive samples.

Reference module: ``torchrec/metrics/num_positive_samples.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import NumPositiveSamplesMetric, NumPositiveSamplesMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_weighted_pos_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (weights.double() * labels.double()).sum(-1)


def get_num_positive_sample_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"weighted_pos_sum": compute_weighted_pos_sum(labels, predictions, weights)}
