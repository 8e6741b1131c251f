"""Calibration after re-calibrating predictions for negative down-sampling.

Reference module: ``torchrec/metrics/calibration_with_recalibration.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .calibration import CalibrationMetricComputation  # noqa: F401
from .ne_with_recalibration import _recalibrate  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class RecalibratedCalibrationMetricComputation(CalibrationMetricComputation):
    """Calibration of re-calibrated predictions. Parity: calibration_with_recalibration.py:25-100."""

    def __init__(self, *args: Any, recalibration_coefficient: float = 1.0, **kwargs: Any) -> None:
        self._recalibration_coefficient = float(recalibration_coefficient)
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return super()._batch_states(_recalibrate(predictions, self._recalibration_coefficient), labels, weights, **kwargs)

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.RECALIBRATED_CALIBRATION, prefix, get("calibration_num") / (get("calibration_denom") + EPS))]


RecalibratedCalibrationMetric = _make("RecalibratedCalibrationMetric", RecalibratedCalibrationMetricComputation, MetricNamespace.RECALIBRATED_CALIBRATION)
