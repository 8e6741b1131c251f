"""Precision of multi-label predictions at a threshold.

Reference module: ``torchrec/metrics/multi_label_precision.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from typing import Any  # noqa: F401


def compute_multi_label_precision(num_true_positives: torch.Tensor, num_false_positives: torch.Tensor) -> torch.Tensor:
    d = num_true_positives + num_false_positives
    return torch.where(d == 0.0, torch.zeros_like(d), num_true_positives / torch.where(d == 0.0, torch.ones_like(d), d)).double()


def decode_integer_to_labels(tensor: torch.Tensor, num_labels: int) -> torch.Tensor:
    """[n] integers -> [n, num_labels] bits, least significant bit first (bit i = label i)."""
    powers = (2 ** torch.arange(num_labels, device=tensor.device)).to(tensor.dtype)
    return ((tensor.reshape(-1, 1) // powers) % 2).to(torch.int32)


class MultiLabelPrecisionMetricComputation(RecMetricComputation):
    """Per-label precision of a multi-label task. Predictions and labels are integers that encode the label set of an example as
    bits (LSB first: 5 = 0b101 -> labels 0 and 2); precision_i = weighted TP_i / (TP_i + FP_i), reported once per label with the
    label name (``label_names``, default ``label_<i>``) as the description, lifetime + window."""

    def __init__(self, *args: Any, num_labels: int = 1, label_names: Optional[List[str]] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._num_labels = num_labels
        self._label_names = [f"label_{i}" for i in range(num_labels)] if label_names is None else list(label_names)
        assert len(self._label_names) == num_labels, "one name per label"
        for name in ("true_pos_sum", "false_pos_sum"):
            self._add_state(name, torch.zeros(self._n_tasks, num_labels, dtype=torch.double), add_window_state=True, dist_reduce_fx="sum", persistent=True)

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        if predictions is None:
            raise RecMetricException("Inputs 'predictions' should not be None for MultiLabelPrecisionMetricComputation update")
        T, L = self._n_tasks, self._num_labels
        pred = decode_integer_to_labels(predictions.reshape(-1).long(), L).view(T, -1, L).double()
        lab = decode_integer_to_labels(labels.reshape(-1).long(), L).view(T, -1, L).double()
        w = (weights.reshape(T, -1, 1).double() if weights is not None else torch.ones(T, pred.shape[1], 1, dtype=torch.double, device=pred.device))
        n = pred.shape[1]
        for name, v in (("true_pos_sum", (w * pred * lab).sum(1)), ("false_pos_sum", (w * pred * (1.0 - lab)).sum(1))):
            st = getattr(self, name)
            v = v.to(st.device)
            st += v
            self._aggregate_window_state(name, v, n)

    def _compute(self) -> List[MetricComputationReport]:
        reports: List[MetricComputationReport] = []
        scopes = [(MetricPrefix.LIFETIME, self.true_pos_sum, self.false_pos_sum)]
        if self._batch_window_buffers is not None:
            scopes.append((MetricPrefix.WINDOW, self.get_window_state("true_pos_sum"), self.get_window_state("false_pos_sum")))
        for i, name in enumerate(self._label_names):
            for prefix, tp, fp in scopes:
                reports.append(MetricComputationReport(MetricName.MULTI_LABEL_PRECISION, prefix, compute_multi_label_precision(tp[:, i], fp[:, i]), description=name))
        return reports


MultiLabelPrecisionMetric = _make("MultiLabelPrecisionMetric", MultiLabelPrecisionMetricComputation, MetricNamespace.MULTI_LABEL_PRECISION)
