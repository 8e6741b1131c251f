"""Area under the ROC curve over a bounded sample window.

Reference module: ``torchrec/metrics/auc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


GROUPING_KEYS = "grouping_keys"


class AUCMetricComputation(_SampleBufferComputation):
    """Windowed weighted AUC. ``grouped_auc``: additionally the mean of the per-group AUCs (``grouped_auc``), the groups given by
    ``required_inputs['grouping_keys']``; ``apply_bin``: soft labels are binarised at 0.039 first."""

    NAME = MetricName.AUC

    def __init__(self, *args: Any, grouped_auc: bool = False, apply_bin: bool = False, fused_update_limit: int = 0, **kwargs: Any) -> None:
        if grouped_auc and fused_update_limit > 0:
            raise RecMetricException("Grouped AUC and Fused Update Limit cannot be enabled together yet.")
        self._grouped_auc, self._apply_bin = grouped_auc, apply_bin
        self.EXTRA = [GROUPING_KEYS] if grouped_auc else []
        super().__init__(*args, fused_update_limit=fused_update_limit, **kwargs)

    def _labels(self, l: torch.Tensor) -> torch.Tensor:
        return (l >= 0.039).to(l.dtype) if self._apply_bin else l

    def _value(self, p, l, w, extra):
        return _auc_from_samples(p, self._labels(l), w)

    def _compute(self) -> List[MetricComputationReport]:
        reports = super()._compute()
        if self._grouped_auc:
            keys = getattr(self, GROUPING_KEYS)[0]
            reports.append(MetricComputationReport(MetricName.GROUPED_AUC, MetricPrefix.WINDOW,
                                                   compute_auc_per_group(self._n_tasks, self.predictions, self._labels(self.labels), self.weights, keys)))
        return reports


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


class AUCMetric(RecMetric):
    _namespace: MetricNamespace = MetricNamespace.AUC
    _computation_class = AUCMetricComputation

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if kwargs.get("grouped_auc"):
            self._required_inputs.add(GROUPING_KEYS)


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
