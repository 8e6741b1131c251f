"""ThroughputMetric: total examples + lifetime / window QPS (reference torchrec/metrics/throughput.py:35)."""
from __future__ import annotations

import logging
import math
import time
from collections import deque
from typing import Deque, Dict

import torch
import torch.nn as nn

from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix, compose_metric_key

logger = logging.getLogger(__name__)

MAX_WINDOW_TS: int = 2 * 60 * 60
MIN_WINDOW_TS: int = 60


class ThroughputMetric(nn.Module):
    """Counts examples of the whole job (``batch_size * world_size`` per ``update``); counters survive
    checkpoints through state-dict buffers."""

    _namespace: MetricNamespace = MetricNamespace.THROUGHPUT
    _metric_name: MetricName = MetricName.THROUGHPUT

    def __init__(self, *, batch_size: int, world_size: int, window_seconds: int, warmup_steps: int = 100, batch_size_stages=None) -> None:
        super().__init__()
        if window_seconds < 1:
            raise ValueError(f"window_seconds must be at least 1 to give window throughput the minimum time window")
        if warmup_steps < 1:
            raise ValueError("warmup_steps must be at least 1 to give throughput a reasonable begin time.")
        if window_seconds > MAX_WINDOW_TS:
            logger.warning(f"window_seconds is greater than {MAX_WINDOW_TS}, capping to {MAX_WINDOW_TS} to make sure window_qps is not staled")
            window_seconds = MAX_WINDOW_TS
        self._batch_size = batch_size
        self._world_size = world_size
        self._window_seconds = window_seconds
        self._warmup_steps = warmup_steps
        self._batch_size_stages = batch_size_stages
        self.register_buffer("total_examples", torch.tensor(0, dtype=torch.long))
        self.register_buffer("warmup_examples", torch.tensor(0, dtype=torch.long))
        self.register_buffer("time_lapse_after_warmup", torch.tensor(0, dtype=torch.double))
        self.register_buffer("attempt_examples", torch.tensor(0, dtype=torch.long), persistent=False)
        self._window_time_lapse_buffer: Deque[float] = deque(maxlen=MAX_WINDOW_TS)
        self._window_time_lapse = 0.0
        self._previous_ts = 0.0
        self._lifetime_key = compose_metric_key(self._namespace, str(self._namespace), self._metric_name, MetricPrefix.LIFETIME)
        self._window_key = compose_metric_key(self._namespace, str(self._namespace), self._metric_name, MetricPrefix.WINDOW)
        self._total_examples_key = compose_metric_key(self._namespace, str(self._namespace), MetricName.TOTAL_EXAMPLES)
        self._attempt_examples_key = compose_metric_key(self._namespace, str(self._namespace), MetricName.ATTEMPT_EXAMPLES)
        self._steps = 0

    def _check_window(self) -> None:
        while self._window_time_lapse > self._window_seconds and len(self._window_time_lapse_buffer) > 1:
            self._window_time_lapse -= self._window_time_lapse_buffer.popleft()

    def update(self) -> None:
        ts = time.monotonic()
        self._steps += 1
        examples = self._batch_size * self._world_size
        self.total_examples += examples
        self.attempt_examples += examples
        if self._steps <= self._warmup_steps:
            self.warmup_examples += examples
            if self._steps == self._warmup_steps:
                self._previous_ts = ts
        else:
            time_lapse = ts - self._previous_ts
            self.time_lapse_after_warmup += time_lapse
            self._window_time_lapse += time_lapse
            self._window_time_lapse_buffer.append(time_lapse)
            self._check_window()
            self._previous_ts = ts

    def compute(self) -> Dict[str, torch.Tensor]:
        ret = {self._total_examples_key: self.total_examples, self._attempt_examples_key: self.attempt_examples}
        if self._steps > self._warmup_steps and (not math.isclose(self.time_lapse_after_warmup.item(), 0) or not math.isclose(self._window_time_lapse, 0)):
            lifetime = (self.total_examples - self.warmup_examples) / self.time_lapse_after_warmup
            if not math.isclose(self._window_time_lapse, 0):
                window = len(self._window_time_lapse_buffer) * self._batch_size * self._world_size / self._window_time_lapse
            else:
                window = 0.0
            if not math.isclose(lifetime.item(), 0):
                ret.update({self._lifetime_key: torch.tensor(lifetime, dtype=torch.double), self._window_key: torch.tensor(window, dtype=torch.double)})
        return ret
