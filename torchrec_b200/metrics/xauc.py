"""Cross AUC for regression: weighted fraction of pairs ordered like their labels.

Reference module: ``torchrec/metrics/xauc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class XAUCMetricComputation(_SumStatesComputation):
    """Cross AUC for regression: the weighted share of the pairs (i < j) of a batch whose predictions are ordered like their labels (a
    pair tied in both counts as ordered), pair weight w_i * w_j; additive states ``error_sum`` / ``weighted_num_pairs``."""

    STATES = ["error_sum", "weighted_num_pairs"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        if predictions is None or weights is None:
            raise RecMetricException("Inputs 'predictions' and 'weights' should not be None for XAUCMetricComputation update")
        return get_xauc_states(labels, predictions, weights)

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.XAUC, prefix, compute_xauc(get("error_sum"), get("weighted_num_pairs")))]


def compute_error_sum(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """Weighted count of pairs (i < j) whose prediction order agrees with their label order."""
    out = []
    for t in range(predictions.shape[0]):
        p, y, w = predictions[t].double(), labels[t].double(), weights[t].double()
        agree = (torch.sign(p[:, None] - p[None, :]) == torch.sign(y[:, None] - y[None, :])).double()
        out.append(torch.triu(agree * (w[:, None] * w[None, :]), diagonal=1).sum())
    return torch.stack(out)


def compute_weighted_num_pairs(weights: torch.Tensor) -> torch.Tensor:
    w = weights.double()
    return torch.stack([torch.triu(w[t][:, None] * w[t][None, :], diagonal=1).sum() for t in range(w.shape[0])])


def compute_xauc(error_sum: torch.Tensor, weighted_num_pairs: torch.Tensor) -> torch.Tensor:
    return torch.where(weighted_num_pairs == 0.0, torch.zeros_like(error_sum), error_sum / weighted_num_pairs).double()


def get_xauc_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"error_sum": compute_error_sum(labels, predictions, weights), "weighted_num_pairs": compute_weighted_num_pairs(weights)}


XAUCMetric = _make("XAUCMetric", XAUCMetricComputation, MetricNamespace.XAUC)
