"""Area under the precision-recall curve.

Reference module: ``torchrec/metrics/auprc.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import AUPRCMetric, AUPRCMetricComputation, _auprc_from_samples  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_auprc(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return torch.stack([_auprc_from_samples(predictions[t], labels[t], weights[t]) for t in range(n_tasks)])


def compute_auprc_per_group(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, grouping_keys: torch.Tensor) -> torch.Tensor:
    out = []
    for t in range(n_tasks):
        vals = [_auprc_from_samples(predictions[t][grouping_keys == g], labels[t][grouping_keys == g], weights[t][grouping_keys == g]) for g in torch.unique(grouping_keys)]
        out.append(torch.stack(vals).mean() if vals else torch.tensor(0.0, dtype=torch.double))
    return torch.stack(out)
