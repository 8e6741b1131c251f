"""``CPUOffloadedRecMetricModule``: metric updates / computes run on a background thread over CPU copies of the model outputs, so the training stream never waits for metrics (reference ``torchrec/metrics/cpu_offloaded_metric_module.py:136``)."""
from __future__ import annotations

import abc
import concurrent.futures
import logging
import queue
import threading
import time
from typing import Any, Dict, List, Optional, Type, Union
import torch
import torch.distributed as dist
import torch.nn as nn
from . import metrics_impl as M
from .metrics_config import MetricsConfig, RecMetricDef, RecMetricEnum, RecMetricEnumBase, StateMetricEnum
from .metrics_namespace import MetricNamespace, compose_customized_metric_key
from .rec_metric import RecMetric, RecMetricList, RecTaskInfo
from .throughput import ThroughputMetric
from .metric_module import MetricValue, RecMetricModule  # noqa: F401


class CPUOffloadedRecMetricModule(RecMetricModule):
    """Asynchronous variant: ``update`` enqueues non-blocking D2H copies of the model outputs and a background
    thread updates / computes the metrics on the CPU, keeping the training stream free
    (reference cpu_offloaded_metric_module.py:136)."""

    def __init__(self, *args: Any, update_queue_size: int = 100, compute_queue_size: int = 100, device: Optional[torch.device] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._queue: "queue.Queue" = queue.Queue(maxsize=update_queue_size)
        self._shutdown = threading.Event()
        self._lock = threading.Lock()
        self._worker = threading.Thread(target=self._run, daemon=True, name="metric_update")
        self._executor = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="metric_compute")
        self._worker.start()

    def update(self, model_out: Dict[str, torch.Tensor], **kwargs: Any) -> None:
        cpu_out: Dict[str, torch.Tensor] = {}
        ev = None
        for k, v in model_out.items():
            if isinstance(v, torch.Tensor) and v.is_cuda:
                buf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                buf.copy_(v.detach(), non_blocking=True)
                cpu_out[k] = buf
            else:
                cpu_out[k] = v
        if torch.cuda.is_available():
            ev = torch.cuda.Event()
            ev.record()
        self.trained_batches += 1
        if self.throughput_metric:
            self.throughput_metric.update()
        self._queue.put((cpu_out, kwargs, ev))

    def _run(self) -> None:
        while not self._shutdown.is_set():
            try:
                item = self._queue.get(timeout=0.1)
            except queue.Empty:
                continue
            out, kwargs, ev = item
            if ev is not None:
                ev.synchronize()
            with self._lock, torch.no_grad():
                self._update_rec_metrics(out, **kwargs)
            self._queue.task_done()

    def compute(self) -> Dict[str, MetricValue]:
        self._queue.join()
        with self._lock:
            return super().compute()

    def async_compute(self) -> "concurrent.futures.Future":
        return self._executor.submit(self.compute)

    def shutdown(self) -> None:
        self._shutdown.set()
        self._worker.join(timeout=5)
        self._executor.shutdown(wait=False)
