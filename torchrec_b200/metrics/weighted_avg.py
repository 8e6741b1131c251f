"""Weighted average of the predictions.

Reference module: ``torchrec/metrics/weighted_avg.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class WeightedAvgMetricComputation(_SumStatesComputation):
    STATES = ["weighted_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"weighted_sum": (predictions.double() * weights.double()).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.WEIGHTED_AVG, prefix, get("weighted_sum") / (get("weighted_num_samples") + EPS))]


WeightedAvgMetric = _make("WeightedAvgMetric", WeightedAvgMetricComputation, MetricNamespace.WEIGHTED_AVG)


def get_mean(value_sum: torch.Tensor, num_samples: torch.Tensor) -> torch.Tensor:
    return value_sum / (num_samples + EPS)
