"""Click-through rate = sum(w * label) / sum(w).

Reference module: ``torchrec/metrics/ctr.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import CTRMetric, CTRMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_ctr(ctr_num: torch.Tensor, ctr_denom: torch.Tensor) -> torch.Tensor:
    return torch.where(ctr_denom == 0.0, torch.zeros_like(ctr_num), ctr_num / ctr_denom).double()


def get_ctr_states(labels: torch.Tensor, predictions: Optional[torch.Tensor], weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"ctr_num": (labels.double() * weights.double()).sum(-1), "ctr_denom": weights.double().sum(-1)}
