"""Calibration of the serving-time prediction.

Reference module: ``torchrec/metrics/serving_calibration.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import ServingCalibrationMetric, ServingCalibrationMetricComputation  # noqa: F401
