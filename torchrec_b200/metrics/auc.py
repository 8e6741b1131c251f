"""Area under the ROC curve over a bounded sample window.

Reference module: ``torchrec/metrics/auc.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import AUCMetric, AUCMetricComputation, _auc_from_samples  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_auc(n_tasks: int, predictions: List[torch.Tensor], labels: List[torch.Tensor], weights: List[torch.Tensor], apply_bin: bool = False) -> torch.Tensor:
    """Weighted AUC per task from buffered samples (lists are concatenated along the sample dim)."""
    p, y, w = (torch.cat(list(x), dim=-1) if isinstance(x, (list, tuple)) else x for x in (predictions, labels, weights))
    return torch.stack([_auc_from_samples(p[t], y[t], w[t]) for t in range(n_tasks)])


def compute_auc_per_group(n_tasks: int, predictions: List[torch.Tensor], labels: List[torch.Tensor], weights: List[torch.Tensor], grouping_keys: torch.Tensor) -> torch.Tensor:
    """Mean of the per-group AUCs (groups = equal values of ``grouping_keys``)."""
    p, y, w = (torch.cat(list(x), dim=-1) if isinstance(x, (list, tuple)) else x for x in (predictions, labels, weights))
    out = []
    for t in range(n_tasks):
        vals = [_auc_from_samples(p[t][grouping_keys == g], y[t][grouping_keys == g], w[t][grouping_keys == g]) for g in torch.unique(grouping_keys)]
        out.append(torch.stack(vals).mean() if vals else torch.tensor(0.5, dtype=torch.double))
    return torch.stack(out)
