"""Precision / recall at the threshold that reaches a target precision in hindsight.

Reference module: ``torchrec/metrics/hindsight_target_pr.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class HindsightTargetPRMetricComputation(RecMetricComputation):
    """Precision/recall at the threshold that reaches a target precision in hindsight (bucketed thresholds)."""

    def __init__(self, *args: Any, target_precision: float = 0.5, threshold_granularity: int = 1000, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._target = target_precision
        self._gran = threshold_granularity
        for s in ["true_pos_sum", "false_pos_sum", "false_neg_sum"]:
            self._add_state(s, torch.zeros(self._n_tasks, threshold_granularity, dtype=torch.double), add_window_state=True, dist_reduce_fx="sum", persistent=True)

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        if predictions is None:
            raise RecMetricException("Inputs 'predictions' should not be None for HindsightTargetPRMetricComputation update")
        if weights is None:
            weights = torch.ones_like(predictions)
        th = torch.linspace(0, 1, self._gran, dtype=torch.double, device=predictions.device).view(1, -1, 1)
        pred = (predictions.double().unsqueeze(1) >= th).double()
        l, w = labels.double().unsqueeze(1), weights.double().unsqueeze(1)
        n = predictions.shape[-1]
        for name, v in (("true_pos_sum", (w * pred * l).sum(-1)), ("false_pos_sum", (w * pred * (1 - l)).sum(-1)), ("false_neg_sum", (w * (1 - pred) * l).sum(-1))):
            st = getattr(self, name)
            v = v.to(st.device)
            st += v
            self._aggregate_window_state(name, v, n)

    def _one(self, tp: torch.Tensor, fp: torch.Tensor, fn: torch.Tensor, prefix: MetricPrefix) -> List[MetricComputationReport]:
        prec = compute_precision(tp, fp)
        ok = prec >= self._target
        idx = torch.where(ok.any(-1), ok.double().argmax(-1), torch.full((self._n_tasks,), self._gran - 1, device=tp.device))  # first threshold bucket reaching the target
        ar = torch.arange(self._n_tasks, device=tp.device)
        return [MetricComputationReport(MetricName.HINDSIGHT_TARGET_PR, prefix, idx.double()),
                MetricComputationReport(MetricName.HINDSIGHT_TARGET_PRECISION, prefix, prec[ar, idx]),
                MetricComputationReport(MetricName.HINDSIGHT_TARGET_RECALL, prefix, compute_recall(tp[ar, idx], fn[ar, idx]))]

    def _compute(self) -> List[MetricComputationReport]:
        out = self._one(self.true_pos_sum, self.false_pos_sum, self.false_neg_sum, MetricPrefix.LIFETIME)
        if self._batch_window_buffers is not None:
            out += self._one(self.get_window_state("true_pos_sum"), self.get_window_state("false_pos_sum"), self.get_window_state("false_neg_sum"), MetricPrefix.WINDOW)
        return out


HindsightTargetPRMetric = _make("HindsightTargetPRMetric", HindsightTargetPRMetricComputation, MetricNamespace.HINDSIGHT_TARGET_PR)


def compute_precision(num_true_positives: torch.Tensor, num_false_positives: torch.Tensor) -> torch.Tensor:
    d = num_true_positives + num_false_positives
    return torch.where(d == 0.0, torch.zeros_like(d), num_true_positives / d).double()


def compute_recall(num_true_positives: torch.Tensor, num_false_negitives: torch.Tensor) -> torch.Tensor:  # (sic: the reference's spelling of the keyword)
    d = num_true_positives + num_false_negitives
    return torch.where(d == 0.0, torch.zeros_like(d), num_true_positives / d).double()


def compute_threshold_idx(num_true_positives: torch.Tensor, num_false_positives: torch.Tensor, target_precision: float) -> int:
    """Smallest threshold bucket whose precision reaches the target (last bucket when none does)."""
    ok = (compute_precision(num_true_positives, num_false_positives) >= target_precision).nonzero()
    return int(ok[0]) if ok.numel() else int(num_true_positives.numel() - 1)


def _bucketed(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, thresholds: torch.Tensor, pos_pred: bool, pos_label: bool) -> torch.Tensor:
    pred = (predictions.double().unsqueeze(-1) >= thresholds.double()) == pos_pred
    lab = ((labels.double() >= 0.5) == pos_label).unsqueeze(-1)
    return (weights.double().unsqueeze(-1) * (pred & lab).double()).sum(-2)


def compute_true_pos_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, thresholds: torch.Tensor) -> torch.Tensor:
    return _bucketed(labels, predictions, weights, thresholds, True, True)


def compute_false_pos_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, thresholds: torch.Tensor) -> torch.Tensor:
    return _bucketed(labels, predictions, weights, thresholds, True, False)


def compute_false_neg_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, thresholds: torch.Tensor) -> torch.Tensor:
    return _bucketed(labels, predictions, weights, thresholds, False, True)


def get_pr_states(labels: torch.Tensor, predictions: torch.Tensor, weights: Optional[torch.Tensor], thresholds: torch.Tensor) -> Dict[str, torch.Tensor]:
    if weights is None:
        weights = torch.ones_like(predictions)
    return {"true_pos_sum": compute_true_pos_sum(labels, predictions, weights, thresholds), "false_pos_sum": compute_false_pos_sum(labels, predictions, weights, thresholds),
            "false_neg_sum": compute_false_neg_sum(labels, predictions, weights, thresholds)}
