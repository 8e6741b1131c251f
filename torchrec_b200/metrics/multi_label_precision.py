"""Precision of multi-label predictions at a threshold.

Reference module: ``torchrec/metrics/multi_label_precision.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import MultiLabelPrecisionMetric, MultiLabelPrecisionMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_multi_label_precision(true_pos_sum: torch.Tensor, false_pos_sum: torch.Tensor) -> torch.Tensor:
    d = true_pos_sum + false_pos_sum
    return torch.where(d == 0.0, torch.zeros_like(d), true_pos_sum / d).double()
