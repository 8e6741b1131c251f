"""Session-level recall of the top-ranked items.

Reference module: ``torchrec/metrics/recall_session.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class RecallSessionMetricComputation(_SampleBufferComputation):
    """Session-level recall: per session, top-k predictions count as positive predictions."""

    NAME = MetricName.RECALL_SESSION_LEVEL
    EXTRA = ["session_ids"]

    def __init__(self, *args: Any, session_metric_def: Optional[Any] = None, top_threshold: int = 1, **kwargs: Any) -> None:
        self._top = getattr(session_metric_def, "top_threshold", None) or top_threshold
        super().__init__(*args, **kwargs)

    def _counts(self, p, l, s):
        tp = fn = fp = 0.0
        for sid in torch.unique(s):
            m = s == sid
            order = torch.argsort(p[m], descending=True)
            pred_pos = torch.zeros(int(m.sum()), dtype=torch.bool)
            pred_pos[order[: self._top]] = True
            lab = l[m] > 0
            tp += float((pred_pos & lab).sum())
            fn += float((~pred_pos & lab).sum())
            fp += float((pred_pos & ~lab).sum())
        return tp, fn, fp

    def _value(self, p, l, w, extra):
        tp, fn, fp = self._counts(p, l, extra["session_ids"])
        return torch.tensor(tp / (tp + fn) if tp + fn > 0 else 0.0, dtype=torch.double)


RecallSessionMetric = _make("RecallSessionMetric", RecallSessionMetricComputation, MetricNamespace.RECALL_SESSION_LEVEL)
