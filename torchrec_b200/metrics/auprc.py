"""Area under the precision-recall curve.

Reference module: ``torchrec/metrics/auprc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class AUPRCMetricComputation(_SampleBufferComputation):
    NAME = MetricName.AUPRC

    def _value(self, p, l, w, extra):
        return _auprc_from_samples(p, l, w)


def _auprc_from_samples(preds, labels, weights) -> torch.Tensor:
    if preds.numel() == 0:
        return torch.tensor(0.0, dtype=torch.double)
    order = torch.argsort(preds, descending=True)
    p, l, w = preds[order].double(), labels[order].double(), weights[order].double()
    ctp = torch.cumsum(w * l, 0)
    cfp = torch.cumsum(w * (1 - l), 0)
    distinct = torch.ones_like(p, dtype=torch.bool)
    distinct[:-1] = p[1:] != p[:-1]
    ctp, cfp = ctp[distinct], cfp[distinct]
    if ctp[-1] == 0:
        return torch.tensor(0.0, dtype=torch.double)
    precision = ctp / (ctp + cfp + EPS)
    recall = ctp / ctp[-1]
    recall_prev = torch.cat([recall.new_zeros(1), recall[:-1]])
    return ((recall - recall_prev) * precision).sum()


AUPRCMetric = _make("AUPRCMetric", AUPRCMetricComputation, MetricNamespace.AUPRC)


def compute_auprc(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return torch.stack([_auprc_from_samples(predictions[t], labels[t], weights[t]) for t in range(n_tasks)])


def compute_auprc_per_group(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, grouping_keys: torch.Tensor) -> torch.Tensor:
    out = []
    for t in range(n_tasks):
        vals = [_auprc_from_samples(predictions[t][grouping_keys == g], labels[t][grouping_keys == g], weights[t][grouping_keys == g]) for g in torch.unique(grouping_keys)]
        out.append(torch.stack(vals).mean() if vals else torch.tensor(0.0, dtype=torch.double))
    return torch.stack(out)
