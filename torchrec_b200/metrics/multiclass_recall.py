"""Recall@k of multi-class predictions.

Reference module: ``torchrec/metrics/multiclass_recall.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make, _zeros  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class MulticlassRecallMetricComputation(RecMetricComputation):
    """recall@k for k in 1..number_of_classes; predictions [n_tasks, B, C]."""

    def __init__(self, *args: Any, number_of_classes: int = 2, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._number_of_classes = number_of_classes
        self._add_state("tp_at_k", torch.zeros(self._n_tasks, number_of_classes, dtype=torch.double), add_window_state=True, dist_reduce_fx="sum")
        self._add_state("total_weights", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="sum")

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        C = self._number_of_classes
        p = predictions.reshape(self._n_tasks, -1, C)
        ranks = torch.argsort(p, dim=-1, descending=True)
        l = labels.reshape(self._n_tasks, -1).long()
        w = weights.reshape(self._n_tasks, -1).double()
        hit_pos = (ranks == l.unsqueeze(-1)).double().argmax(-1)  # rank position of the true class
        tp = torch.zeros(self._n_tasks, C, dtype=torch.double)
        for k in range(C):
            tp[:, k] = (w * (hit_pos <= k).double()).sum(-1)
        self.tp_at_k += tp.to(self.tp_at_k.device)
        self.total_weights += w.sum(-1).to(self.total_weights.device)
        self._aggregate_window_state("tp_at_k", tp, l.shape[-1])
        self._aggregate_window_state("total_weights", w.sum(-1), l.shape[-1])

    def _compute(self) -> List[MetricComputationReport]:
        out = [MetricComputationReport(MetricName.MULTICLASS_RECALL, MetricPrefix.LIFETIME, self.tp_at_k / (self.total_weights.unsqueeze(-1) + EPS))]
        if self._batch_window_buffers is not None:
            out.append(MetricComputationReport(MetricName.MULTICLASS_RECALL, MetricPrefix.WINDOW,
                                               self.get_window_state("tp_at_k") / (self.get_window_state("total_weights").unsqueeze(-1) + EPS)))
        return out


MulticlassRecallMetric = _make("MulticlassRecallMetric", MulticlassRecallMetricComputation, MetricNamespace.MULTICLASS_RECALL)


def compute_true_positives_at_k(predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, n_classes: int) -> torch.Tensor:
    """``[n_classes]``: entry k-1 = weighted number of samples whose label is among the top-k classes."""
    ranks = torch.argsort(predictions, dim=-1, descending=True)
    hit_at = (ranks == labels.long().unsqueeze(-1)).double()  # [N, C] one-hot of the label's rank
    return (torch.cumsum(hit_at, dim=-1) * weights.double().unsqueeze(-1)).sum(0)


def compute_multiclass_recall_at_k(tp_at_k: torch.Tensor, total_weights: torch.Tensor) -> torch.Tensor:
    return tp_at_k / (total_weights.unsqueeze(-1) + EPS) if tp_at_k.dim() > total_weights.dim() else tp_at_k / (total_weights + EPS)


def get_multiclass_recall_states(predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, n_classes: int) -> Dict[str, torch.Tensor]:
    return {"tp_at_k": compute_true_positives_at_k(predictions, labels, weights, n_classes), "total_weights": weights.double().sum(-1)}
