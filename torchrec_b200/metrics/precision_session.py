"""Session-level precision of the top-ranked items.

Reference module: ``torchrec/metrics/precision_session.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .recall_session import RecallSessionMetricComputation  # noqa: F401


class PrecisionSessionMetricComputation(RecallSessionMetricComputation):
    NAME = MetricName.PRECISION_SESSION_LEVEL

    def _value(self, p, l, w, extra):
        tp, fn, fp = self._counts(p, l, extra["session_ids"])
        return torch.tensor(tp / (tp + fp) if tp + fp > 0 else 0.0, dtype=torch.double)


PrecisionSessionMetric = _make("PrecisionSessionMetric", PrecisionSessionMetricComputation, MetricNamespace.PRECISION_SESSION_LEVEL)
