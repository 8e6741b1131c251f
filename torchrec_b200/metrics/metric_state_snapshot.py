"""Point-in-time copy of every metric's state tensors (reference metrics/metric_state_snapshot.py:24-125): what an asynchronous compute
works on while the training thread keeps updating the live metrics."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .rec_metric import RecMetricList
from .throughput import ThroughputMetric


class MetricStateSnapshot:
    def __init__(self, metric_states: Dict[str, Dict[str, torch.Tensor]], throughput_metric: Optional[ThroughputMetric] = None) -> None:
        self.metric_states = metric_states            # "<metric namespace>_<task idx>" -> {state name: tensor copy}
        self.throughput_metric = throughput_metric

    @classmethod
    def from_metrics(cls, rec_metrics: RecMetricList, throughput_metric: Optional[ThroughputMetric] = None, device: str = "cpu") -> "MetricStateSnapshot":
        states: Dict[str, Dict[str, torch.Tensor]] = {}
        for metric in rec_metrics.rec_metrics:
            ns = getattr(metric._namespace, "value", str(metric._namespace))
            for i, comp in enumerate(metric._metrics_computations):
                states[f"{ns}_{i}"] = {n: comp._buffers[n].detach().to(device, copy=True) for n in comp._state_names}
        return cls(states, throughput_metric)

    def load_into(self, rec_metrics: RecMetricList) -> None:
        """Write the snapshot into (a clone of) the metric list the compute thread owns."""
        with torch.no_grad():
            for metric in rec_metrics.rec_metrics:
                ns = getattr(metric._namespace, "value", str(metric._namespace))
                for i, comp in enumerate(metric._metrics_computations):
                    saved = self.metric_states.get(f"{ns}_{i}")
                    if saved is None:
                        continue
                    for n, t in saved.items():
                        cur = comp._buffers[n]
                        if cur.shape == t.shape:
                            cur.copy_(t)
                        else:  # sample-buffer states grow: replace
                            comp._buffers[n] = t.to(cur.device).clone()
