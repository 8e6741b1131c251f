"""MSE normalised by the MSE of a constant-one predictor.

Reference module: ``torchrec/metrics/nmse.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class NMSEMetricComputation(_SumStatesComputation):
    """MSE normalised by the MSE of the constant predictor 1."""

    STATES = ["error_sum", "const_pred_error_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        d = predictions.double() - labels.double()
        c = 1.0 - labels.double()
        return {"error_sum": (weights.double() * d * d).sum(-1), "const_pred_error_sum": (weights.double() * c * c).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        nmse = get("error_sum") / (get("const_pred_error_sum") + EPS)
        return [MetricComputationReport(MetricName.NMSE, prefix, nmse), MetricComputationReport(MetricName.NRMSE, prefix, torch.sqrt(nmse))]


NMSEMetric = _make("NMSEMetric", NMSEMetricComputation, MetricNamespace.NMSE)


def compute_norm(model_error_sum: torch.Tensor, baseline_error_sum: torch.Tensor) -> torch.Tensor:
    return torch.where(baseline_error_sum == 0.0, torch.zeros_like(model_error_sum), model_error_sum / baseline_error_sum).double()


def get_norm_mse_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    w, y = weights.double(), labels.double()
    return {"error_sum": (w * (y - predictions.double()) ** 2).sum(-1), "weighted_num_samples": w.sum(-1), "const_pred_error_sum": (w * (y - 1.0) ** 2).sum(-1)}
