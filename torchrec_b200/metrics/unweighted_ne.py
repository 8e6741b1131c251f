"""NE with unit weights.

Reference module: ``torchrec/metrics/unweighted_ne.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .ne import NEMetricComputation  # noqa: F401


class UnweightedNEMetricComputation(NEMetricComputation):
    def _batch_states(self, predictions, labels, weights, **kwargs):
        return super()._batch_states(predictions, labels, torch.ones_like(weights), **kwargs)

    def _reports(self, get, prefix):
        reps = super()._reports(get, prefix)
        reps[0] = MetricComputationReport(MetricName.UNWEIGHTED_NE, prefix, reps[0].value)
        return reps


UnweightedNEMetric = _make("UnweightedNEMetric", UnweightedNEMetricComputation, MetricNamespace.UNWEIGHTED_NE)
