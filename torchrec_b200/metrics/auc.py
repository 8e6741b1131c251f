"""Area under the ROC curve over a bounded sample window.

Reference module: ``torchrec/metrics/auc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class AUCMetricComputation(_SampleBufferComputation):
    NAME = MetricName.AUC

    def _value(self, p, l, w, extra):
        return _auc_from_samples(p, l, w)


# ---- sample-buffer metrics --------------------------------------------------------------------------------------
def _auc_from_samples(preds: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """Weighted ROC AUC with tie handling (trapezoid over the sorted-by-score cumulative TP/FP curve)."""
    if preds.numel() == 0:
        return torch.tensor(0.5, dtype=torch.double)
    order = torch.argsort(preds, descending=True)
    p, l, w = preds[order].double(), labels[order].double(), weights[order].double()
    ctp = torch.cumsum(w * l, 0)
    cfp = torch.cumsum(w * (1 - l), 0)
    # keep only the last point of every group of tied scores
    distinct = torch.ones_like(p, dtype=torch.bool)
    distinct[:-1] = p[1:] != p[:-1]
    ctp, cfp = ctp[distinct], cfp[distinct]
    ctp = torch.cat([ctp.new_zeros(1), ctp])
    cfp = torch.cat([cfp.new_zeros(1), cfp])
    if ctp[-1] == 0 or cfp[-1] == 0:
        return torch.tensor(0.5, dtype=torch.double)
    return torch.trapz(ctp, cfp) / (ctp[-1] * cfp[-1])


AUCMetric = _make("AUCMetric", AUCMetricComputation, MetricNamespace.AUC)


def compute_auc(n_tasks: int, predictions: List[torch.Tensor], labels: List[torch.Tensor], weights: List[torch.Tensor], apply_bin: bool = False) -> torch.Tensor:
    """Weighted AUC per task from buffered samples (lists are concatenated along the sample dim)."""
    p, y, w = (torch.cat(list(x), dim=-1) if isinstance(x, (list, tuple)) else x for x in (predictions, labels, weights))
    return torch.stack([_auc_from_samples(p[t], y[t], w[t]) for t in range(n_tasks)])


def compute_auc_per_group(n_tasks: int, predictions: List[torch.Tensor], labels: List[torch.Tensor], weights: List[torch.Tensor], grouping_keys: torch.Tensor) -> torch.Tensor:
    """Mean of the per-group AUCs (groups = equal values of ``grouping_keys``)."""
    p, y, w = (torch.cat(list(x), dim=-1) if isinstance(x, (list, tuple)) else x for x in (predictions, labels, weights))
    out = []
    for t in range(n_tasks):
        vals = [_auc_from_samples(p[t][grouping_keys == g], y[t][grouping_keys == g], w[t][grouping_keys == g]) for g in torch.unique(grouping_keys)]
        out.append(torch.stack(vals).mean() if vals else torch.tensor(0.5, dtype=torch.double))
    return torch.stack(out)
