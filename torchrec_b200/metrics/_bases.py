"""Shared bases of the metric implementations: additive per-task sum states (``_SumStatesComputation``, ``_SingleSumComputation``), bounded
sample buffers for rank-based metrics (``_SampleBufferComputation``) and the ``RecMetric`` class factory. Every metric lives in its own module
(``ne.py``, ``auc.py``, ...) on top of these."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Type

import torch

from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException

EPS = torch.finfo(torch.float64).eps


def _zeros(n: int) -> torch.Tensor:
    return torch.zeros(n, dtype=torch.double)


class _SumStatesComputation(RecMetricComputation):
    """Base for metrics whose states are per-task sums: subclasses define STATES and ``_batch_states`` / ``_value``."""

    STATES: List[str] = []

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        for s in self.STATES:
            self._add_state(s, _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="sum", persistent=True)

    def _batch_states(self, predictions, labels, weights, **kwargs) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    def _reports(self, get) -> List[MetricComputationReport]:
        raise NotImplementedError

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        if predictions is None and "predictions" in self._needs():
            raise RecMetricException(f"Inputs 'predictions' should not be None for {type(self).__name__} update")
        states = self._batch_states(predictions, labels, weights, **kwargs)
        n = labels.shape[-1]
        for name, v in states.items():
            st = getattr(self, name)
            v = v.to(st.dtype).to(st.device)
            st += v
            self._aggregate_window_state(name, v, n)

    def _needs(self) -> List[str]:
        return ["predictions"]

    def _compute(self) -> List[MetricComputationReport]:
        reports = self._reports(lambda n: getattr(self, n), MetricPrefix.LIFETIME)
        if self._batch_window_buffers is not None:
            reports += self._reports(lambda n: self.get_window_state(n), MetricPrefix.WINDOW)
        return reports + self._extra_reports()

    def _extra_reports(self) -> List[MetricComputationReport]:
        """Reports outside the lifetime / window pair (e.g. example counters)."""
        return []


class _SampleBufferComputation(RecMetricComputation):
    """Keeps the last ``window_size`` (pred, label, weight[, group]) samples per task; all-gathered at compute."""

    EXTRA: List[str] = []

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        for n in ["predictions", "labels", "weights"] + self.EXTRA:
            self._add_state(n, torch.zeros(self._n_tasks, 0, dtype=torch.double), add_window_state=False, dist_reduce_fx="cat", persistent=False)

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        if predictions is None or weights is None:
            raise RecMetricException(f"Inputs 'predictions' and 'weights' should not be None for {type(self).__name__} update")
        vals = {"predictions": predictions, "labels": labels, "weights": weights}
        for e in self.EXTRA:
            v = kwargs.get(e)
            if v is None and "required_inputs" in kwargs:
                v = kwargs["required_inputs"].get(e)
            if v is None:
                raise RecMetricException(f"{type(self).__name__} needs '{e}'")
            vals[e] = v.reshape(1, -1).expand(self._n_tasks, -1) if v.dim() == 1 else v
        cap = self._window_size
        for n, v in vals.items():
            cur = getattr(self, n)
            new = torch.cat([cur, v.reshape(self._n_tasks, -1).double().to(cur.device)], dim=-1)
            if new.shape[-1] > cap:
                new = new[:, -cap:]
            self._buffers[n] = new

    def _value(self, p, l, w, extra) -> torch.Tensor:
        raise NotImplementedError

    NAME = MetricName.AUC

    def _compute(self) -> List[MetricComputationReport]:
        vals = []
        for t in range(self._n_tasks):
            extra = {e: getattr(self, e)[t] for e in self.EXTRA}
            vals.append(self._value(self.predictions[t], self.labels[t], self.weights[t], extra))
        return [MetricComputationReport(self.NAME, MetricPrefix.WINDOW, torch.stack(vals))]

    def reset(self) -> None:
        for n in ["predictions", "labels", "weights"] + self.EXTRA:
            self._buffers[n] = torch.zeros(self._n_tasks, 0, dtype=torch.double)


class _SingleSumComputation(_SumStatesComputation):
    """One weighted sum reported as-is (data-volume monitors)."""

    NAME: MetricName

    def _needs(self):
        return []

    def _sum(self, predictions, labels, weights) -> torch.Tensor:
        raise NotImplementedError

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {self.STATES[0]: self._sum(predictions, labels, weights)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(self.NAME, prefix, get(self.STATES[0]))]


def _make(name: str, comp: Type[RecMetricComputation], ns: MetricNamespace) -> Type[RecMetric]:
    return type(name, (RecMetric,), {"_namespace": ns, "_computation_class": comp, "__doc__": comp.__doc__})
