"""NE of the serving-time prediction.

Reference module: ``torchrec/metrics/serving_ne.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .ne import NEMetricComputation  # noqa: F401


class ServingNEMetricComputation(NEMetricComputation):
    def _reports(self, get, prefix):
        reps = super()._reports(get, prefix)
        return [MetricComputationReport(MetricName.SERVING_NE, prefix, reps[0].value)]


ServingNEMetric = _make("ServingNEMetric", ServingNEMetricComputation, MetricNamespace.SERVING_NE)
