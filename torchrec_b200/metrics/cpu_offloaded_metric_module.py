"""``CPUOffloadedRecMetricModule`` under its reference import path (``torchrec/metrics/cpu_offloaded_metric_module.py``); implementation in ``metric_module.py``."""
from .metric_module import CPUOffloadedRecMetricModule  # noqa: F401
