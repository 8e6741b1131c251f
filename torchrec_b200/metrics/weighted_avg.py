"""Weighted average of the predictions.

Reference module: ``torchrec/metrics/weighted_avg.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import WeightedAvgMetric, WeightedAvgMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def get_mean(value_sum: torch.Tensor, num_samples: torch.Tensor) -> torch.Tensor:
    return value_sum / (num_samples + EPS)
