"""MSE normalised by the MSE of a constant-one predictor.

Reference module: ``torchrec/metrics/nmse.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import NMSEMetric, NMSEMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_norm(model_error_sum: torch.Tensor, baseline_error_sum: torch.Tensor) -> torch.Tensor:
    return torch.where(baseline_error_sum == 0.0, torch.zeros_like(model_error_sum), model_error_sum / baseline_error_sum).double()


def get_norm_mse_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    return {"error_sum": (w * (y - predictions.double()) ** 2).sum(-1), "weighted_num_samples": w.sum(-1), "const_pred_error_sum": (w * (y - 1.0) ** 2).sum(-1)}
