"""Metric module config (reference ``torchrec/distributed/test_utils/metric_config.py:36`` ``RecMetricConfig``)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch


@dataclass
class RecMetricConfig:
    metrics: List[str] = field(default_factory=lambda: ["ne", "auc"])
    task_name: str = "task"
    label_name: str = "label"
    prediction_name: str = "prediction"
    weight_name: str = "weight"
    window_size: int = 10_000_000
    throughput: bool = True
    compute_interval_steps: int = 100

    def generate_metric_module(self, batch_size: int, world_size: int, rank: int, device: Optional[torch.device] = None):
        from ...metrics import MetricsConfig, RecMetricDef, RecMetricEnum, RecTaskInfo, ThroughputDef, generate_metric_module
        from ...metrics.metric_module import RecMetricModule

        task = RecTaskInfo(name=self.task_name, label_name=self.label_name, prediction_name=self.prediction_name, weight_name=self.weight_name)
        cfg = MetricsConfig(rec_tasks=[task], rec_metrics={RecMetricEnum(m): RecMetricDef(rec_tasks=[task], window_size=self.window_size) for m in self.metrics},
                            throughput_metric=ThroughputDef() if self.throughput else None, compute_interval_steps=self.compute_interval_steps)
        return generate_metric_module(RecMetricModule, cfg, batch_size=batch_size, world_size=world_size, my_rank=rank, state_metrics_mapping={},
                                      device=device or torch.device("cpu"))
