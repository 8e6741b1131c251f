"""Normalized entropy over positive samples.

Reference module: ``torchrec/metrics/ne_positive.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class NEPositiveMetricComputation(_SumStatesComputation):
    """Normalized entropy of the POSITIVE samples only: ``-sum w*y*log2(p)`` over the base-rate entropy of the positives
    (reference metrics/ne_positive.py:25-70)."""

    STATES = ["cross_entropy_positive_sum", "weighted_num_samples", "pos_labels", "neg_labels"]

    def __init__(self, *args: Any, allow_missing_label_with_zero_weight: bool = False, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.eta = 1e-12

    def _batch_states(self, predictions, labels, weights, **kwargs):
        p = torch.clamp(predictions.double(), self.eta, 1 - self.eta)
        w, y = weights.double(), labels.double()
        return {"cross_entropy_positive_sum": (-w * y * torch.log2(p)).sum(-1), "weighted_num_samples": w.sum(-1),
                "pos_labels": (w * y).sum(-1), "neg_labels": (w * (1 - y)).sum(-1)}

    def _reports(self, get, prefix):
        mean_label = get("pos_labels") / (get("weighted_num_samples") + EPS)
        norm = -get("pos_labels") * torch.log2(mean_label + self.eta)
        return [MetricComputationReport(MetricName.NE_POSITIVE, prefix, get("cross_entropy_positive_sum") / (norm + EPS))]


NEPositiveMetric = _make("NEPositiveMetric", NEPositiveMetricComputation, MetricNamespace.NE_POSITIVE)


def compute_cross_entropy_positive(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> torch.Tensor:
    return -weights.double() * labels.double() * torch.log2(torch.clamp(predictions.double(), eta, 1 - eta))


def compute_ne_positive(ce_positive_sum: torch.Tensor, weighted_num_samples: torch.Tensor, pos_labels: torch.Tensor, neg_labels: torch.Tensor, eta: float,
                        allow_missing_label_with_zero_weight: bool = False) -> torch.Tensor:
    mean_label = pos_labels / (weighted_num_samples + EPS)
    ne = ce_positive_sum / (-pos_labels * torch.log2(mean_label + eta) + EPS)
    if allow_missing_label_with_zero_weight and not bool(torch.as_tensor(weighted_num_samples).all()):
        return torch.where(weighted_num_samples > 0, ne, torch.full_like(ne, eta))
    return ne


def get_ne_positive_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    return {"cross_entropy_positive_sum": compute_cross_entropy_positive(labels, predictions, weights, eta).sum(-1), "weighted_num_samples": w.sum(-1),
            "pos_labels": (w * y).sum(-1), "neg_labels": (w * (1.0 - y)).sum(-1)}
