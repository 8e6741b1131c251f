"""NE after removing the calibration error.

Reference module: ``torchrec/metrics/cali_free_ne.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .ne import NEMetricComputation, compute_ne  # noqa: F401


class CaliFreeNEMetricComputation(NEMetricComputation):
    """Calibration-free NE: NE divided by the cross entropy the mean prediction would have on the observed label counts,
    ``-pos * log2(mean_pred) - (N - pos) * log2(1 - mean_pred)`` - a constant rescaling of all predictions barely moves it."""

    STATES = NEMetricComputation.STATES + ["weighted_sum_predictions"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        st = super()._batch_states(predictions, labels, weights, **kwargs)
        st["weighted_sum_predictions"] = (weights.double() * predictions.double()).sum(-1)
        return st

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.CALI_FREE_NE, prefix, compute_cali_free_ne(get("cross_entropy_sum"), get("weighted_num_samples"), get("pos_labels"),
                                                                                         get("neg_labels"), get("weighted_sum_predictions"), self.eta))]


def compute_cali_free_ne(ce_sum: torch.Tensor, weighted_num_samples: torch.Tensor, pos_labels: torch.Tensor, neg_labels: torch.Tensor,
                         weighted_sum_predictions: torch.Tensor, eta: float, allow_missing_label_with_zero_weight: bool = False) -> torch.Tensor:
    if allow_missing_label_with_zero_weight and not bool(weighted_num_samples.all()):
        return torch.tensor([eta], dtype=torch.double)
    n = weighted_num_samples.double().clamp(min=eta)
    mean_label = (pos_labels / n).clamp(min=eta, max=1 - eta)
    raw_ne = ce_sum / (-pos_labels * torch.log2(mean_label) - neg_labels * torch.log2(1.0 - mean_label))
    mean_pred = weighted_sum_predictions / weighted_num_samples
    return raw_ne / (-pos_labels * torch.log2(mean_pred) - (weighted_num_samples - pos_labels) * torch.log2(1.0 - mean_pred))


def get_cali_free_ne_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    p = predictions.double().clamp(min=eta, max=1 - eta)
    ce = -w * y * torch.log2(p) - w * (1.0 - y) * torch.log2(1.0 - p)
    return {"cross_entropy_sum": ce.sum(-1), "weighted_num_samples": w.sum(-1), "pos_labels": (w * y).sum(-1), "neg_labels": (w * (1.0 - y)).sum(-1),
            "weighted_sum_predictions": (w * predictions.double()).sum(-1)}


CaliFreeNEMetric = _make("CaliFreeNEMetric", CaliFreeNEMetricComputation, MetricNamespace.CALI_FREE_NE)


def compute_cross_entropy(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> torch.Tensor:
    p = predictions.double().clamp(min=eta, max=1 - eta)
    return -weights.double() * labels.double() * torch.log2(p) - weights.double() * (1.0 - labels.double()) * torch.log2(1.0 - p)
