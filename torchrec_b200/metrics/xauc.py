"""Cross AUC for regression: weighted fraction of pairs ordered like their labels.

Reference module: ``torchrec/metrics/xauc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class XAUCMetricComputation(_SampleBufferComputation):
    """Cross AUC for regression: P(pred_i > pred_j | label_i > label_j), weighted by w_i * w_j."""

    NAME = MetricName.XAUC

    def _value(self, p, l, w, extra):
        n = p.numel()
        if n < 2:
            return torch.tensor(0.0, dtype=torch.double)
        if n > 4096:
            idx = torch.randperm(n)[:4096]
            p, l, w = p[idx], l[idx], w[idx]
        ww = w.unsqueeze(0) * w.unsqueeze(1)
        dp = torch.sign(p.unsqueeze(0) - p.unsqueeze(1))
        dl = torch.sign(l.unsqueeze(0) - l.unsqueeze(1))
        match = ((dp == dl) & (dl != 0)).double()
        iu = torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1)
        return (ww * match)[iu].sum() / (ww[iu].sum() + EPS)


XAUCMetric = _make("XAUCMetric", XAUCMetricComputation, MetricNamespace.XAUC)


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
