"""Auto-encoder reconstruction loss observed at serving time (additive state).

Reference module: ``torchrec/metrics/serving_ae_loss.py``."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class ServingAELossMetricComputation(_SumStatesComputation):
    """Weighted mean absolute error of the served prediction (the reference only reserves the name: metrics_config.py:48)."""

    STATES = ["error_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"error_sum": (weights.double() * (labels.double() - predictions.double()).abs()).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.SERVING_AE_LOSS, prefix, get("error_sum") / (get("weighted_num_samples") + EPS))]


ServingAELossMetric = _make("ServingAELossMetric", ServingAELossMetricComputation, MetricNamespace.SERVING_AE_LOSS)
