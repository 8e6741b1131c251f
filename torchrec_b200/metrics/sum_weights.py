"""Sum of the sample weights.

Reference module: ``torchrec/metrics/sum_weights.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import SumWeightsMetric, SumWeightsMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_weighted_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return weights.double().sum(-1)


def get_weighted_sum_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"weighted_sum": compute_weighted_sum(labels, predictions, weights)}
