"""Precision of multi-label predictions at a threshold.

Reference module: ``torchrec/metrics/multi_label_precision.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .precision import PrecisionMetricComputation  # noqa: F401


class MultiLabelPrecisionMetricComputation(PrecisionMetricComputation):
    def _reports(self, get, prefix):
        r = super()._reports(get, prefix)
        return [MetricComputationReport(MetricName.MULTI_LABEL_PRECISION, prefix, r[0].value)]


MultiLabelPrecisionMetric = _make("MultiLabelPrecisionMetric", MultiLabelPrecisionMetricComputation, MetricNamespace.MULTI_LABEL_PRECISION)


def compute_multi_label_precision(true_pos_sum: torch.Tensor, false_pos_sum: torch.Tensor) -> torch.Tensor:
    d = true_pos_sum + false_pos_sum
    return torch.where(d == 0.0, torch.zeros_like(d), true_pos_sum / d).double()
