"""Weighted number of positive samples.

Reference module: ``torchrec/metrics/num_positive_samples.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SingleSumComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class NumPositiveSamplesMetricComputation(_SingleSumComputation):
    """sum w * label (NaN labels count 0). Parity: num_positive_samples.py:21-95."""

    STATES = ["weighted_pos_sum"]
    NAME = MetricName.NUM_POSITIVE_SAMPLES

    def _sum(self, predictions, labels, weights):
        return (weights.double() * torch.nan_to_num(labels.double(), 0.0)).sum(-1)


NumPositiveSamplesMetric = _make("NumPositiveSamplesMetric", NumPositiveSamplesMetricComputation, MetricNamespace.NUM_POSITIVE_SAMPLES)


def compute_weighted_pos_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return (weights.double() * labels.double()).sum(-1)


def get_num_positive_sample_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"weighted_pos_sum": compute_weighted_pos_sum(labels, predictions, weights)}
