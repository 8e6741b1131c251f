"""NE per segment of a grouping key.

Reference module: ``torchrec/metrics/segmented_ne.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .ne import _ce, compute_ne  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class SegmentedNEMetricComputation(RecMetricComputation):
    """NE per segment (``grouping_keys`` in [0, num_groups))."""

    def __init__(self, *args: Any, num_groups: int = 1, grouping_keys: str = "grouping_keys", **kwargs: Any) -> None:
        kwargs.pop("include_logloss", None)
        super().__init__(*args, **kwargs)
        self._num_groups = num_groups
        for s in ["cross_entropy_sum", "weighted_num_samples", "pos_labels", "neg_labels"]:
            self._add_state(s, torch.zeros(self._n_tasks, num_groups, dtype=torch.double), add_window_state=False, dist_reduce_fx="sum")

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        g = kwargs.get("grouping_keys")
        if g is None and "required_inputs" in kwargs:
            g = kwargs["required_inputs"].get("grouping_keys")
        g = g.reshape(1, -1).expand(self._n_tasks, -1).long()
        ce = _ce(labels, predictions, weights)
        w = weights.double()
        for name, v in (("cross_entropy_sum", ce), ("weighted_num_samples", w), ("pos_labels", w * labels.double()), ("neg_labels", w * (1 - labels.double()))):
            st = getattr(self, name)
            st.scatter_add_(1, g.to(st.device), v.to(st.device))

    def _compute(self) -> List[MetricComputationReport]:
        ne = compute_ne(self.cross_entropy_sum, self.weighted_num_samples, self.pos_labels, self.neg_labels)
        return [MetricComputationReport(MetricName.SEGMENTED_NE, MetricPrefix.LIFETIME, ne[:, gi], description=f"_{gi}") for gi in range(self._num_groups)]


SegmentedNEMetric = _make("SegmentedNEMetric", SegmentedNEMetricComputation, MetricNamespace.SEGMENTED_NE)
