"""Calibration = sum(w * prediction) / sum(w * label).

Reference module: ``torchrec/metrics/calibration.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class CalibrationMetricComputation(_SumStatesComputation):
    STATES = ["calibration_num", "calibration_denom"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"calibration_num": (predictions.double() * weights.double()).sum(-1), "calibration_denom": (labels.double() * weights.double()).sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.CALIBRATION, prefix, get("calibration_num") / (get("calibration_denom") + EPS))]


CalibrationMetric = _make("CalibrationMetric", CalibrationMetricComputation, MetricNamespace.CALIBRATION)


def compute_calibration(calibration_num: torch.Tensor, calibration_denom: torch.Tensor) -> torch.Tensor:
    return torch.where(calibration_denom <= 0.0, torch.zeros_like(calibration_num), calibration_num / calibration_denom).double()


def get_calibration_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"calibration_num": (predictions.double() * weights.double()).sum(-1), "calibration_denom": (labels.double() * weights.double()).sum(-1)}
