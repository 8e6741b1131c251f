"""Click-through rate = sum(w * label) / sum(w).

Reference module: ``torchrec/metrics/ctr.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class CTRMetricComputation(_SumStatesComputation):
    STATES = ["ctr_num", "ctr_denom"]

    def _needs(self):
        return []

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"ctr_num": (labels.double() * weights.double()).sum(-1), "ctr_denom": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.CTR, prefix, get("ctr_num") / (get("ctr_denom") + EPS))]


CTRMetric = _make("CTRMetric", CTRMetricComputation, MetricNamespace.CTR)


def compute_ctr(ctr_num: torch.Tensor, ctr_denom: torch.Tensor) -> torch.Tensor:
    return torch.where(ctr_denom == 0.0, torch.zeros_like(ctr_num), ctr_num / ctr_denom).double()


def get_ctr_states(labels: torch.Tensor, predictions: Optional[torch.Tensor], weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"ctr_num": (labels.double() * weights.double()).sum(-1), "ctr_denom": weights.double().sum(-1)}
