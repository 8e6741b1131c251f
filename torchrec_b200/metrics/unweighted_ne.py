"""NE with unit weights.

Reference module: ``torchrec/metrics/unweighted_ne.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .ne import NEMetricComputation  # noqa: F401


class UnweightedNEMetricComputation(NEMetricComputation):
    def _batch_states(self, predictions, labels, weights, **kwargs):
        return super()._batch_states(predictions, labels, torch.ones_like(weights), **kwargs)

    def _reports(self, get, prefix):
        reps = super()._reports(get, prefix)
        reps[0] = MetricComputationReport(MetricName.UNWEIGHTED_NE, prefix, reps[0].value)
        return reps


UnweightedNEMetric = _make("UnweightedNEMetric", UnweightedNEMetricComputation, MetricNamespace.UNWEIGHTED_NE)


def compute_cross_entropy(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> torch.Tensor:
    """Base-2 cross entropy per example with every weight taken as 1."""
    p = torch.clamp(predictions.double(), eta, 1 - eta)
    y = labels.double()
    return -(y * torch.log2(p) + (1.0 - y) * torch.log2(1.0 - p))


def compute_ne(ce_sum: torch.Tensor, weighted_num_samples: torch.Tensor, pos_labels: torch.Tensor, neg_labels: torch.Tensor, eta: float,
               allow_missing_label_with_zero_weight: bool = False) -> torch.Tensor:
    from .ne import compute_ne as _compute_ne

    return _compute_ne(ce_sum, weighted_num_samples, pos_labels, neg_labels, eta, allow_missing_label_with_zero_weight)


def get_unweighted_ne_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, eta: float) -> Dict[str, torch.Tensor]:
    y = labels.double()
    ones = torch.ones_like(y)
    return {"cross_entropy_sum": compute_cross_entropy(labels, predictions, weights, eta).sum(-1), "weighted_num_samples": ones.sum(-1), "pos_labels": y.sum(-1),
            "neg_labels": (1.0 - y).sum(-1)}
