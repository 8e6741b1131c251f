"""RecMetric framework: per-task windowed + lifetime metric states with cross-rank aggregation
(reference torchrec/metrics/rec_metric.py:127-900).

A ``RecMetricComputation`` keeps additive states (sums) for one metric; ``RecMetric`` owns one computation per
task (or one fused computation over all tasks stacked on dim 0), a sliding window of the last
``window_size`` examples, and ``compute()`` returns ``{key: value}`` with lifetime and window variants.
States are plain tensors that are summed across ranks with one all-reduce per state at compute time."""
from __future__ import annotations

import abc
from collections import deque
from dataclasses import dataclass
from enum import Enum
from typing import Any, Callable, Deque, Dict, Iterator, List, Optional, Tuple, Type, Union

import torch
import torch.distributed as dist
from torch import nn

from .metrics_namespace import MetricName, MetricNamespaceBase, MetricPrefix, compose_metric_key


class RecComputeMode(Enum):
    FUSED_TASKS_COMPUTATION = 1
    UNFUSED_TASKS_COMPUTATION = 2
    FUSED_TASKS_AND_STATES_COMPUTATION = 3


RecModelOutput = Union[torch.Tensor, Dict[str, torch.Tensor]]


@dataclass
class RecTaskInfo:
    name: str = "DefaultTask"
    label_name: str = "label"
    prediction_name: str = "prediction"
    weight_name: str = "weight"
    session_metric_def: Optional[Any] = None
    is_negative_task: bool = False
    tensor_name: Optional[str] = None
    weighted: bool = True


@dataclass
class MetricComputationReport:
    name: MetricName
    metric_prefix: MetricPrefix
    value: torch.Tensor
    description: Optional[str] = None


DefaultValueT = Union[torch.Tensor, List[torch.Tensor]]
ComputeIterType = Iterator[Tuple[RecTaskInfo, MetricName, torch.Tensor, str]]
MAX_BUFFER_COUNT = 1000


class RecMetricException(Exception):
    pass


class WindowBuffer:
    """Sliding window over per-batch state contributions: keeps the batches covering the last ``max_size`` examples."""

    def __init__(self, max_size: int, max_buffer_count: int) -> None:
        self._max_size = max_size
        self._max_buffer_count = max_buffer_count
        self._buffers: Deque[torch.Tensor] = deque(maxlen=max_buffer_count)
        self._used_sizes: Deque[int] = deque(maxlen=max_buffer_count)
        self._window_used_size = 0

    def aggregate_state(self, window_state: torch.Tensor, curr_state: torch.Tensor, size: int) -> None:
        def remove(window_state: torch.Tensor) -> None:
            window_state -= self._buffers.popleft()
            self._window_used_size -= self._used_sizes.popleft()

        if len(self._buffers) == self._buffers.maxlen:
            remove(window_state)
        self._buffers.append(curr_state)
        self._used_sizes.append(size)
        window_state += curr_state
        self._window_used_size += size
        while self._window_used_size > self._max_size and len(self._buffers) > 1:
            remove(window_state)

    @property
    def buffers(self) -> Deque[torch.Tensor]:
        return self._buffers


class RecMetricComputation(nn.Module, abc.ABC):
    """One metric's additive state for ``n_tasks`` stacked tasks. Subclasses call ``_add_state`` in ``__init__`` and
    implement ``update`` / ``_compute``."""

    _batch_window_buffers: Optional[Dict[str, WindowBuffer]]

    def __init__(self, my_rank: int, batch_size: int, n_tasks: int, window_size: int, compute_on_all_ranks: bool = False,
                 should_validate_update: bool = False, fuse_state_tensors: bool = False, process_group: Optional[dist.ProcessGroup] = None,
                 fused_update_limit: int = 0, allow_missing_label_with_zero_weight: bool = False, *args: Any, **kwargs: Any) -> None:
        super().__init__()
        self._my_rank = my_rank
        self._n_tasks = n_tasks
        self._batch_size = batch_size
        self._window_size = window_size
        self._compute_on_all_ranks = compute_on_all_ranks
        self._should_validate_update = should_validate_update
        self._process_group = process_group
        self._state_names: List[str] = []
        self._reductions: Dict[str, str] = {}
        self._persistent: Dict[str, bool] = {}
        self._batch_window_buffers = {} if window_size > 0 else None
        self._allow_missing_label_with_zero_weight = allow_missing_label_with_zero_weight

    @staticmethod
    def get_window_state_name(state_name: str) -> str:
        return f"window_{state_name}"

    def get_window_state(self, state_name: str) -> torch.Tensor:
        return getattr(self, self.get_window_state_name(state_name))

    def _add_state(self, name: str, default: DefaultValueT, add_window_state: bool, dist_reduce_fx: str = "sum", persistent: bool = True, **kwargs: Any) -> None:
        self.register_buffer(name, default.clone() if isinstance(default, torch.Tensor) else default, persistent=persistent)
        self._state_names.append(name)
        self._reductions[name] = dist_reduce_fx
        if add_window_state and self._batch_window_buffers is not None:
            wname = self.get_window_state_name(name)
            self.register_buffer(wname, default.clone(), persistent=False)
            self._state_names.append(wname)
            self._reductions[wname] = dist_reduce_fx
            self._batch_window_buffers[wname] = WindowBuffer(max_size=self._window_size, max_buffer_count=MAX_BUFFER_COUNT)

    def _aggregate_window_state(self, state_name: str, state: torch.Tensor, num_samples: int) -> None:
        if self._batch_window_buffers is None:
            return
        wname = self.get_window_state_name(state_name)
        self._batch_window_buffers[wname].aggregate_state(getattr(self, wname), curr_state=state.detach().clone(), size=num_samples)

    @abc.abstractmethod
    def update(self, *, predictions: Optional[torch.Tensor], labels: torch.Tensor, weights: Optional[torch.Tensor], **kwargs: Any) -> None:
        ...

    @abc.abstractmethod
    def _compute(self) -> List[MetricComputationReport]:
        ...

    def pre_compute(self) -> None:
        return

    def synced_states(self) -> Dict[str, torch.Tensor]:
        """States reduced over the process group (sum / max / cat)."""
        out: Dict[str, torch.Tensor] = {}
        pg = self._process_group
        distributed = dist.is_initialized() and (pg is not None or dist.get_world_size() > 1)
        for n in self._state_names:
            t = getattr(self, n)
            if isinstance(t, list):
                t = torch.cat(t, dim=-1) if t else torch.zeros(self._n_tasks, 0)
            red = self._reductions[n]
            if distributed:
                if red == "sum":
                    t = t.clone()
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pg)
                elif red == "max":
                    t = t.clone()
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=pg)
                elif red == "mean":
                    t = t.clone()
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pg)
                    t /= dist.get_world_size(pg)
                elif red == "cat":
                    W = dist.get_world_size(pg)
                    sizes = [torch.zeros(1, dtype=torch.long, device=t.device) for _ in range(W)]
                    dist.all_gather(sizes, torch.tensor([t.shape[-1]], device=t.device), group=pg)
                    mx = int(max(int(s) for s in sizes))
                    pad = torch.zeros(t.shape[:-1] + (mx,), dtype=t.dtype, device=t.device)
                    pad[..., : t.shape[-1]] = t
                    gathered = [torch.zeros_like(pad) for _ in range(W)]
                    dist.all_gather(gathered, pad, group=pg)
                    t = torch.cat([g[..., : int(s)] for g, s in zip(gathered, sizes)], dim=-1)
            out[n] = t
        return out

    def compute(self) -> List[MetricComputationReport]:
        """Reduce the states over the process group, then evaluate (on rank 0 unless ``compute_on_all_ranks``)."""
        self.pre_compute()
        saved = {n: self._buffers[n] for n in self._state_names}
        try:
            for n, t in self.synced_states().items():
                self._buffers[n] = t
            if self._my_rank == 0 or self._compute_on_all_ranks:
                return self._compute()
            return []
        finally:
            for n, t in saved.items():
                self._buffers[n] = t

    def local_compute(self) -> List[MetricComputationReport]:
        return self._compute()

    def reset(self) -> None:
        for n in self._state_names:
            t = getattr(self, n)
            if isinstance(t, torch.Tensor):
                t.zero_()
        if self._batch_window_buffers is not None:
            for k in list(self._batch_window_buffers.keys()):
                self._batch_window_buffers[k] = WindowBuffer(self._window_size, MAX_BUFFER_COUNT)


class RecMetric(nn.Module, abc.ABC):
    """Metric over one or several tasks; see module docstring."""

    _computation_class: Type[RecMetricComputation]
    _namespace: MetricNamespaceBase

    def __init__(self, world_size: int, my_rank: int, batch_size: int, tasks: List[RecTaskInfo], compute_mode: RecComputeMode = RecComputeMode.UNFUSED_TASKS_COMPUTATION,
                 window_size: int = 100, fused_update_limit: int = 0, compute_on_all_ranks: bool = False, should_validate_update: bool = False,
                 process_group: Optional[dist.ProcessGroup] = None, **kwargs: Any) -> None:
        super().__init__()
        torch._C._log_api_usage_once(f"torchrec_b200.metrics.rec_metric.{self.__class__.__name__}")
        if window_size < batch_size:
            raise ValueError(f"Local window size must be larger than batch size. Got local window size {window_size} and batch size {batch_size}.")
        self._world_size = world_size
        self._my_rank = my_rank
        self._window_size = window_size * world_size  # global window measured in examples seen by all ranks
        self._batch_size = batch_size
        self._metrics_computations: nn.ModuleList = nn.ModuleList()
        self._tasks = tasks
        self._compute_mode = compute_mode
        self._fused_update_limit = fused_update_limit
        self._should_validate_update = should_validate_update
        self._default_weight: Dict[Tuple[int, ...], torch.Tensor] = {}
        self._required_inputs: set = set()
        self._update_buffers: Dict[str, List[Any]] = {}
        self._check_fused_update_limit()
        if self._compute_mode in (RecComputeMode.FUSED_TASKS_COMPUTATION, RecComputeMode.FUSED_TASKS_AND_STATES_COMPUTATION):
            task_per_metric = len(self._tasks)
            self._tasks_iter = self._fused_tasks_iter
        else:
            task_per_metric = 1
            self._tasks_iter = self._unfused_tasks_iter
        for task_config in tasks if task_per_metric == 1 else [tasks]:
            kwargs["fused_update_limit"] = fused_update_limit
            kw = dict(kwargs)
            if task_per_metric == 1 and getattr(task_config, "session_metric_def", None) is not None:
                kw["session_metric_def"] = task_config.session_metric_def
            kw.update(self._get_task_kwargs(task_config))
            self._required_inputs.update(self._get_task_required_inputs(task_config))
            self._metrics_computations.append(self._computation_class(
                my_rank, batch_size, task_per_metric, self._window_size, compute_on_all_ranks,
                should_validate_update, process_group=process_group, **kw))

    def _get_task_kwargs(self, task_config: Union[RecTaskInfo, List[RecTaskInfo]]) -> Dict[str, Any]:
        """Extra constructor arguments of the computation of one task (of all tasks when they are fused)."""
        return {}

    def _get_task_required_inputs(self, task_config: Union[RecTaskInfo, List[RecTaskInfo]]) -> set:
        """Names of the extra model outputs the computation of this task reads from ``required_inputs``."""
        return set()

    def _check_fused_update_limit(self) -> None:
        if self._fused_update_limit > 0 and self._compute_mode == RecComputeMode.UNFUSED_TASKS_COMPUTATION:
            pass

    def _fused_tasks_iter(self, compute_scope: str) -> ComputeIterType:
        assert len(self._metrics_computations) == 1
        for report in getattr(self._metrics_computations[0], compute_scope)():
            for task, value in zip(self._tasks, report.value if report.value.dim() > 0 else report.value.view(1)):
                yield task, report.name, value.view(-1), compute_scope + report.metric_prefix.value, report.description  # type: ignore[misc]

    def _unfused_tasks_iter(self, compute_scope: str) -> ComputeIterType:
        for task, metric_computation in zip(self._tasks, self._metrics_computations):
            for report in getattr(metric_computation, compute_scope)():
                yield task, report.name, report.value, compute_scope + report.metric_prefix.value, report.description  # type: ignore[misc]

    def _create_default_weights(self, predictions: torch.Tensor) -> torch.Tensor:
        w = self._default_weight.get(tuple(predictions.size()))
        if w is None or w.device != predictions.device:
            w = torch.ones_like(predictions, dtype=torch.float32)
            self._default_weight[tuple(predictions.size())] = w
        return w

    def _check_nonempty_weights(self, weights: torch.Tensor) -> torch.Tensor:
        return torch.gt(torch.count_nonzero(weights, dim=-1), 0)

    def update(self, *, predictions, labels, weights=None, **kwargs: Any) -> None:
        """``predictions/labels/weights``: dict task-name -> tensor (or stacked tensors when tasks are fused)."""
        with torch.no_grad():
            if self._compute_mode in (RecComputeMode.FUSED_TASKS_COMPUTATION, RecComputeMode.FUSED_TASKS_AND_STATES_COMPUTATION):
                if isinstance(predictions, dict):
                    p = torch.stack([predictions[t.name].view(-1) for t in self._tasks]) if predictions is not None else None
                    l = torch.stack([labels[t.name].view(-1) for t in self._tasks])
                    w = torch.stack([weights[t.name].view(-1) for t in self._tasks]) if weights is not None else None
                else:
                    p, l, w = predictions, labels, weights
                if w is None and p is not None:
                    w = self._create_default_weights(p)
                req = kwargs.get("required_inputs")
                if req is not None:  # per-task tensors (TensorWeightedAvg) travel stacked in task order
                    targets = [req[t.tensor_name] for t in self._tasks if t.tensor_name and t.tensor_name in req]
                    if targets:
                        kwargs = dict(kwargs)
                        kwargs["required_inputs"] = dict(req)
                        kwargs["required_inputs"]["target_tensor"] = torch.stack([x.reshape(-1) for x in targets]).view(len(targets), -1)
                self._metrics_computations[0].update(predictions=p, labels=l, weights=w, **kwargs)
            else:
                for task, metric_ in zip(self._tasks, self._metrics_computations):
                    if predictions is not None and task.name not in predictions:
                        continue
                    p = predictions[task.name].view(1, -1) if predictions is not None else None
                    l = labels[task.name].view(1, -1)
                    if weights is None or task.name not in weights:
                        w = self._create_default_weights(l.float())
                    else:
                        w = weights[task.name].view(1, -1)
                    if w.numel() and not bool(self._check_nonempty_weights(w).any()):
                        continue  # all-zero weights: nothing to add
                    task_kwargs = {k: (v[task.name] if isinstance(v, dict) and task.name in v else v) for k, v in kwargs.items()}
                    metric_.update(predictions=p, labels=l, weights=w, **task_kwargs)

    def compute(self) -> Dict[str, torch.Tensor]:
        return {compose_metric_key(self._namespace, task.name, metric_name, _prefix(metric_prefix), description): value
                for task, metric_name, value, metric_prefix, description in self._tasks_iter("compute")}

    def local_compute(self) -> Dict[str, torch.Tensor]:
        return {compose_metric_key(self._namespace, task.name, metric_name, _prefix(metric_prefix), description): value
                for task, metric_name, value, metric_prefix, description in self._tasks_iter("local_compute")}

    def reset(self) -> None:
        for computation in self._metrics_computations:
            computation.reset()

    def get_memory_usage(self) -> Dict[torch.Tensor, int]:
        out: Dict[torch.Tensor, int] = {}
        for c in self._metrics_computations:
            for b in c.buffers():
                out[b] = b.numel() * b.element_size()
        return out

    def get_required_inputs(self) -> set:
        return self._required_inputs


def _prefix(scope_and_prefix: str) -> MetricPrefix:
    for p in (MetricPrefix.LIFETIME, MetricPrefix.WINDOW, MetricPrefix.ATTEMPT):
        if scope_and_prefix.endswith(p.value):
            return p
    return MetricPrefix.DEFAULT


class RecMetricList(nn.Module):
    """List of RecMetrics behind the RecMetric interface."""

    rec_metrics: nn.ModuleList
    required_inputs: Optional[List[str]]

    def __init__(self, rec_metrics: List[RecMetric]) -> None:
        super().__init__()
        self.rec_metrics = nn.ModuleList(rec_metrics)
        self.required_inputs = list(set().union(*[m.get_required_inputs() for m in rec_metrics])) or None if rec_metrics else None

    def __len__(self) -> int:
        return len(self.rec_metrics)

    def __getitem__(self, idx: int) -> nn.Module:
        return self.rec_metrics[idx]

    def get_required_inputs(self) -> Optional[List[str]]:
        return self.required_inputs

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        for metric in self.rec_metrics:
            metric.update(predictions=predictions, labels=labels, weights=weights, **kwargs)

    def compute(self) -> Dict[str, torch.Tensor]:
        ret = {}
        for metric in self.rec_metrics:
            ret.update(metric.compute())
        return ret

    def local_compute(self) -> Dict[str, torch.Tensor]:
        ret = {}
        for metric in self.rec_metrics:
            ret.update(metric.local_compute())
        return ret

    def reset(self) -> None:
        for metric in self.rec_metrics:
            metric.reset()
