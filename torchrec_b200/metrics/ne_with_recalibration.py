"""NE after re-calibrating predictions for negative down-sampling.

Reference module: ``torchrec/metrics/ne_with_recalibration.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .ne import NEMetricComputation  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class RecalibratedNEMetricComputation(NEMetricComputation):
    """NE on predictions re-calibrated for the training-time negative down-sampling rate. Parity: ne_with_recalibration.py:20-116."""

    def __init__(self, *args: Any, recalibration_coefficient: float = 1.0, **kwargs: Any) -> None:
        self._recalibration_coefficient = float(recalibration_coefficient)
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return super()._batch_states(_recalibrate(predictions, self._recalibration_coefficient), labels, weights, **kwargs)

    def _reports(self, get, prefix):
        reps = super()._reports(get, prefix)
        return [MetricComputationReport(MetricName.RECALIBRATED_NE, prefix, reps[0].value)] + reps[1:]


def _recalibrate(predictions: torch.Tensor, coef: float) -> torch.Tensor:
    """Undo negative down-sampling: p -> p / (p + (1 - p) / c). Parity: ne_with_recalibration.py:76-85."""
    p = predictions.double()
    return p / (p + (1.0 - p) / coef)


RecalibratedNEMetric = _make("RecalibratedNEMetric", RecalibratedNEMetricComputation, MetricNamespace.RECALIBRATED_NE)
