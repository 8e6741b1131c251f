"""Session-level normalised discounted cumulative gain.

Reference module: ``torchrec/metrics/ndcg.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class NDCGMetricComputation(_SampleBufferComputation):
    """Session NDCG: samples grouped by ``session_ids``; gain = label (or 2^label - 1 with exponential_gain)."""

    NAME = MetricName.NDCG
    EXTRA = ["session_ids"]

    def __init__(self, *args: Any, exponential_gain: bool = False, k: int = -1, **kwargs: Any) -> None:
        kwargs.pop("session_key", None)
        self._exp = exponential_gain
        self._k = k
        super().__init__(*args, **kwargs)

    def _value(self, p, l, w, extra):
        s = extra["session_ids"]
        vals = []
        for sid in torch.unique(s):
            m = s == sid
            gains = (2.0 ** l[m] - 1.0) if self._exp else l[m]
            k = gains.numel() if self._k <= 0 else min(self._k, gains.numel())
            disc = 1.0 / torch.log2(torch.arange(2, k + 2, dtype=torch.double))
            dcg = (gains[torch.argsort(p[m], descending=True)][:k] * disc).sum()
            idcg = (torch.sort(gains, descending=True).values[:k] * disc).sum()
            if idcg > 0:
                vals.append(dcg / idcg)
        return torch.stack(vals).mean() if vals else torch.tensor(0.0, dtype=torch.double)


NDCGMetric = _make("NDCGMetric", NDCGMetricComputation, MetricNamespace.NDCG)


NDCGComputation = NDCGMetricComputation
