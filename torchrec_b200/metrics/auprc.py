"""Area under the precision-recall curve.

Reference module: ``torchrec/metrics/auprc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


GROUPING_KEYS = "grouping_keys"


class AUPRCMetricComputation(_SampleBufferComputation):
    """Windowed weighted area under the precision-recall curve; ``grouped_auprc`` adds the mean over the groups of
    ``required_inputs['grouping_keys']`` (``grouped_auprc``)."""

    NAME = MetricName.AUPRC

    def __init__(self, *args: Any, grouped_auprc: bool = False, fused_update_limit: int = 0, **kwargs: Any) -> None:
        if grouped_auprc and fused_update_limit > 0:
            raise RecMetricException("Grouped AUPRC and Fused Update Limit cannot be enabled together yet.")
        self._grouped_auprc = grouped_auprc
        self.EXTRA = [GROUPING_KEYS] if grouped_auprc else []
        super().__init__(*args, fused_update_limit=fused_update_limit, **kwargs)

    def _value(self, p, l, w, extra):
        return _auprc_from_samples(p, l, w)

    def _compute(self) -> List[MetricComputationReport]:
        reports = super()._compute()
        if self._grouped_auprc:
            reports.append(MetricComputationReport(MetricName.GROUPED_AUPRC, MetricPrefix.WINDOW,
                                                   compute_auprc_per_group(self._n_tasks, self.predictions, self.labels, self.weights, getattr(self, GROUPING_KEYS)[0])))
        return reports


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


class AUPRCMetric(RecMetric):
    _namespace: MetricNamespace = MetricNamespace.AUPRC
    _computation_class = AUPRCMetricComputation

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if kwargs.get("grouped_auprc"):
            self._required_inputs.add(GROUPING_KEYS)


def compute_auprc(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return torch.stack([_auprc_from_samples(predictions[t], labels[t], weights[t]) for t in range(n_tasks)])


def compute_auprc_per_group(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, grouping_keys: torch.Tensor) -> torch.Tensor:
    out = []
    for t in range(n_tasks):
        vals = [_auprc_from_samples(predictions[t][grouping_keys == g], labels[t][grouping_keys == g], weights[t][grouping_keys == g]) for g in torch.unique(grouping_keys)]
        out.append(torch.stack(vals).mean() if vals else torch.tensor(0.0, dtype=torch.double))
    return torch.stack(out)
