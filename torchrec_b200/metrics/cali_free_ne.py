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
    """NE after rescaling predictions to perfect calibration (isolates ranking quality)."""

    STATES = NEMetricComputation.STATES + ["weighted_sum_predictions"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        st = super()._batch_states(predictions, labels, weights, **kwargs)
        st["weighted_sum_predictions"] = (weights.double() * predictions.double()).sum(-1)
        return st

    def _reports(self, get, prefix):
        ne = compute_ne(get("cross_entropy_sum"), get("weighted_num_samples"), get("pos_labels"), get("neg_labels"), self.eta)
        mean_pred = get("weighted_sum_predictions") / (get("weighted_num_samples") + EPS)
        mean_label = get("pos_labels") / (get("weighted_num_samples") + EPS)
        ent = lambda q: -(mean_label * torch.log2(q + self.eta) + (1 - mean_label) * torch.log2(1 - q + self.eta))
        cali_term = (ent(mean_pred) - ent(mean_label)) / (ent(mean_label) + EPS)
        return [MetricComputationReport(MetricName.CALI_FREE_NE, prefix, ne - cali_term)]


CaliFreeNEMetric = _make("CaliFreeNEMetric", CaliFreeNEMetricComputation, MetricNamespace.CALI_FREE_NE)
