"""Count of samples whose label is missing (negative sentinel).

Reference module: ``torchrec/metrics/num_missing_labels.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import NumMissingLabelsMetric, NumMissingLabelsMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_missing_label_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (labels.double() < 0).double().sum(-1)


def get_num_missing_labels_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"missing_label_sum": compute_missing_label_sum(labels, predictions, weights)}
