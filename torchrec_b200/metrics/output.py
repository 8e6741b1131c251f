"""Statistics of the raw model output.

Reference module: ``torchrec/metrics/output.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class OutputMetricComputation(_SumStatesComputation):
    """Mean prediction and mean label (model output monitoring)."""

    STATES = ["latest_imp", "total_latest_imp"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"latest_imp": (predictions.double() * weights.double()).sum(-1), "total_latest_imp": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.OUTPUT, prefix, get("latest_imp") / (get("total_latest_imp") + EPS))]


OutputMetric = _make("OutputMetric", OutputMetricComputation, MetricNamespace.OUTPUT)
