"""Calibration = sum(w * prediction) / sum(w * label).

Reference module: ``torchrec/metrics/calibration.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import CalibrationMetric, CalibrationMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def compute_calibration(calibration_num: torch.Tensor, calibration_denom: torch.Tensor) -> torch.Tensor:
    return torch.where(calibration_denom <= 0.0, torch.zeros_like(calibration_num), calibration_num / calibration_denom).double()


def get_calibration_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"calibration_num": (predictions.double() * weights.double()).sum(-1), "calibration_denom": (labels.double() * weights.double()).sum(-1)}
