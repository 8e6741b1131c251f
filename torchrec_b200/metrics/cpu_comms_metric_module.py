"""Metric module of the compute thread in the CPU-offloaded pipeline (reference metrics/cpu_comms_metric_module.py:26-163): it owns a
CLONE of the metric list, loads state snapshots into it and runs the cross-rank sync + compute over a CPU (gloo) process group, so the
NCCL stream of the trainer never carries metric traffic."""
from __future__ import annotations

import copy
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist

from .metric_module import MetricValue, RecMetricModule
from .metric_state_snapshot import MetricStateSnapshot
from .rec_metric import RecMetricComputation, RecMetricList


def set_update_called(computation: RecMetricComputation) -> None:
    """Loaded states count as "updated": compute must not warn about missing updates."""
    if hasattr(computation, "_update_called"):
        computation._update_called = True  # type: ignore[attr-defined]


class CPUCommsRecMetricModule(RecMetricModule):
    def __init__(self, *args: Any, cpu_process_group: Optional[dist.ProcessGroup] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._cpu_pg = cpu_process_group
        self.rec_metrics = self._clone_rec_metrics()

    def _clone_rec_metrics(self) -> RecMetricList:
        cloned = copy.deepcopy(self.rec_metrics)
        for metric in cloned.rec_metrics:
            for comp in metric._metrics_computations:
                comp.to("cpu")
                if self._cpu_pg is not None and hasattr(comp, "process_group"):
                    comp.process_group = self._cpu_pg
        return cloned

    def load_local_metric_state_snapshot(self, snapshot: MetricStateSnapshot) -> None:
        snapshot.load_into(self.rec_metrics)
        for metric in self.rec_metrics.rec_metrics:
            for comp in metric._metrics_computations:
                set_update_called(comp)
        if snapshot.throughput_metric is not None:
            self.throughput_metric = snapshot.throughput_metric

    def compute_from_snapshot(self, snapshot: MetricStateSnapshot) -> Dict[str, MetricValue]:
        self.load_local_metric_state_snapshot(snapshot)
        return self.compute()
