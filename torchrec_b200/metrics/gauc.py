"""Grouped AUC.

Reference module: ``torchrec/metrics/gauc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .auc import _auc_from_samples  # noqa: F401


class GroupedAUCMetricComputation(_SampleBufferComputation):
    """GAUC: mean of per-group AUCs (groups with a single class are skipped)."""

    NAME = MetricName.GROUPED_AUC
    EXTRA = ["grouping_keys"]

    def _value(self, p, l, w, extra):
        g = extra["grouping_keys"]
        aucs = []
        for k in torch.unique(g):
            m = g == k
            if l[m].min() == l[m].max():
                continue
            aucs.append(_auc_from_samples(p[m], l[m], w[m]))
        return torch.stack(aucs).mean() if aucs else torch.tensor(0.5, dtype=torch.double)


GAUCMetric = _make("GAUCMetric", GroupedAUCMetricComputation, MetricNamespace.GROUPED_AUC)


GAUCMetricComputation = GroupedAUCMetricComputation
