"""Normalized entropy.

Reference module: ``torchrec/metrics/ne.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


def _ce(labels, preds, weights, eta=1e-12):
    p = torch.clamp(preds.double(), eta, 1 - eta)
    return -weights.double() * (labels.double() * torch.log2(p) + (1 - labels.double()) * torch.log2(1 - p))


class NEMetricComputation(_SumStatesComputation):
    """Normalized entropy = cross entropy / entropy of the base rate. ``include_logloss`` adds logloss."""

    STATES = ["cross_entropy_sum", "weighted_num_samples", "pos_labels", "neg_labels"]

    def __init__(self, *args: Any, include_logloss: bool = False, allow_missing_label_with_zero_weight: bool = False, **kwargs: Any) -> None:
        self._include_logloss = include_logloss
        self._allow_missing_label = allow_missing_label_with_zero_weight
        super().__init__(*args, **kwargs)
        self.eta = 1e-12

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"cross_entropy_sum": _ce(labels, predictions, weights, self.eta).sum(-1), "weighted_num_samples": weights.double().sum(-1),
                "pos_labels": (weights.double() * labels.double()).sum(-1), "neg_labels": (weights.double() * (1 - labels.double())).sum(-1)}

    def _reports(self, get, prefix):
        ne = compute_ne(get("cross_entropy_sum"), get("weighted_num_samples"), get("pos_labels"), get("neg_labels"), self.eta, getattr(self, "_allow_missing_label", False))
        out = [MetricComputationReport(MetricName.NE, prefix, ne)]
        if self._include_logloss:
            ll = get("cross_entropy_sum") / (get("weighted_num_samples") + EPS) * torch.log(torch.tensor(2.0, dtype=torch.double))
            out.append(MetricComputationReport(MetricName.LOG_LOSS, prefix, ll))
        return out


def compute_ne(ce_sum, weighted_num_samples, pos_labels, neg_labels, eta=1e-12, allow_missing_label_with_zero_weight: bool = False) -> torch.Tensor:
    """cross entropy / cross entropy of the base rate. ``allow_missing_label_with_zero_weight``: tasks that saw no weight yet report ``eta``
    instead of 0 / 0."""
    mean_label = pos_labels / (weighted_num_samples + EPS)
    ce_norm = -(pos_labels * torch.log2(mean_label + eta) + neg_labels * torch.log2(1 - mean_label + eta))
    ne = ce_sum / (ce_norm + EPS)
    if allow_missing_label_with_zero_weight and not bool(torch.as_tensor(weighted_num_samples).all()):
        return torch.where(weighted_num_samples > 0, ne, torch.full_like(ne, eta))
    return ne


NEMetric = _make("NEMetric", NEMetricComputation, MetricNamespace.NE)


def compute_cross_entropy(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> torch.Tensor:
    p = torch.clamp(predictions.double(), eta, 1 - eta)
    return -weights.double() * (labels.double() * torch.log2(p) + (1.0 - labels.double()) * torch.log2(1.0 - p))


def compute_logloss(ce_sum: torch.Tensor, pos_labels: torch.Tensor, neg_labels: torch.Tensor, eta: float) -> torch.Tensor:
    """Mean natural-log loss from the base-2 cross entropy sum."""
    n = pos_labels + neg_labels
    n = torch.where(n == 0.0, torch.full_like(n, eta), n)
    return torch.log(torch.tensor(2.0, dtype=ce_sum.dtype, device=ce_sum.device)) * ce_sum / n


def get_ne_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    return {"cross_entropy_sum": compute_cross_entropy(labels, predictions, weights, eta).sum(-1), "weighted_num_samples": w.sum(-1),
            "pos_labels": (w * y).sum(-1), "neg_labels": (w * (1.0 - y)).sum(-1)}
