"""RecMetricModule: update / compute entry point used by training loops
(reference torchrec/metrics/metric_module.py:197-900, cpu_offloaded_metric_module.py:136)."""
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

logger = logging.getLogger(__name__)

REC_METRICS_MAPPING: Dict[RecMetricEnumBase, Type[RecMetric]] = {
    RecMetricEnum.NE: M.NEMetric, RecMetricEnum.SEGMENTED_NE: M.SegmentedNEMetric, RecMetricEnum.CTR: M.CTRMetric,
    RecMetricEnum.CALIBRATION: M.CalibrationMetric, RecMetricEnum.AUC: M.AUCMetric, RecMetricEnum.AUPRC: M.AUPRCMetric, RecMetricEnum.RAUC: M.RAUCMetric,
    RecMetricEnum.MSE: M.MSEMetric, RecMetricEnum.MAE: M.MAEMetric, RecMetricEnum.MULTICLASS_RECALL: M.MulticlassRecallMetric,
    RecMetricEnum.WEIGHTED_AVG: M.WeightedAvgMetric, RecMetricEnum.TOWER_QPS: M.TowerQPSMetric, RecMetricEnum.RECALL_SESSION_LEVEL: M.RecallSessionMetric,
    RecMetricEnum.PRECISION_SESSION_LEVEL: M.PrecisionSessionMetric, RecMetricEnum.ACCURACY: M.AccuracyMetric, RecMetricEnum.NDCG: M.NDCGMetric,
    RecMetricEnum.XAUC: M.XAUCMetric, RecMetricEnum.SCALAR: M.ScalarMetric, RecMetricEnum.PRECISION: M.PrecisionMetric, RecMetricEnum.RECALL: M.RecallMetric,
    RecMetricEnum.SERVING_NE: M.ServingNEMetric, RecMetricEnum.SERVING_CALIBRATION: M.ServingCalibrationMetric, RecMetricEnum.OUTPUT: M.OutputMetric,
    RecMetricEnum.TENSOR_WEIGHTED_AVG: M.TensorWeightedAvgMetric, RecMetricEnum.CALI_FREE_NE: M.CaliFreeNEMetric, RecMetricEnum.UNWEIGHTED_NE: M.UnweightedNEMetric,
    RecMetricEnum.HINDSIGHT_TARGET_PR: M.HindsightTargetPRMetric, RecMetricEnum.NMSE: M.NMSEMetric, RecMetricEnum.AVERAGE: M.AverageMetric,
    RecMetricEnum.GAUC: M.GAUCMetric, RecMetricEnum.MULTI_LABEL_PRECISION: M.MultiLabelPrecisionMetric,
    RecMetricEnum.RECALIBRATED_NE: M.RecalibratedNEMetric, RecMetricEnum.RECALIBRATED_CALIBRATION: M.RecalibratedCalibrationMetric,
    RecMetricEnum.SERVING_AE_LOSS: M.ServingAELossMetric, RecMetricEnum.NUM_POSITIVE_SAMPLES: M.NumPositiveSamplesMetric,
    RecMetricEnum.SUM_WEIGHTS: M.SumWeightsMetric, RecMetricEnum.NUM_MISSING_LABELS: M.NumMissingLabelsMetric,
    RecMetricEnum.WEIGHTED_SUM_PREDICTIONS: M.WeightedSumPredictionsMetric, RecMetricEnum.NE_POSITIVE: M.NEPositiveMetric,
}

MODEL_METRIC_LABEL: str = "model_out"
MEMORY_AVG_WARNING_PERCENTAGE = 20
MEMORY_AVG_WARNING_WARMUP = 100

MetricValue = Union[torch.Tensor, float]
MetricsResult = Dict[str, MetricValue]  # what ``compute`` returns
MetricsFuture = concurrent.futures.Future  # of a MetricsResult (asynchronous compute)
MetricsOutput = Union[MetricsResult, MetricsFuture, "DeferrableMetrics"]
PublishableMetrics = Dict[str, Any]  # after the values were made plain (floats / lists) for a metrics sink
PublishableMetricsFuture = concurrent.futures.Future
PublishableMetricsOutput = Union[PublishableMetrics, PublishableMetricsFuture, "DeferrableMetrics"]


class StateMetric(abc.ABC):
    """Metrics computed from model / optimizer state instead of model outputs."""

    @abc.abstractmethod
    def get_metrics(self) -> Dict[str, MetricValue]:
        ...


def parse_task_model_outputs(tasks: List[RecTaskInfo], model_out: Dict[str, torch.Tensor], required_inputs_list: Optional[List[str]] = None):
    """Split a flat ``model_out`` dict into per-task labels / predictions / weights (+ required inputs)."""
    all_labels: Dict[str, torch.Tensor] = {}
    all_predictions: Dict[str, torch.Tensor] = {}
    all_weights: Dict[str, torch.Tensor] = {}
    all_required_inputs: Dict[str, torch.Tensor] = {}
    for task in tasks:
        labels, predictions, weights = model_out[task.label_name], model_out[task.prediction_name], model_out.get(task.weight_name)
        if labels.size() == predictions.size():
            labels, predictions = labels.squeeze(), predictions.squeeze()
        has_valid_weights = weights is None or bool(torch.gt(torch.count_nonzero(weights), 0))
        if predictions.numel() > 0 and has_valid_weights:
            all_labels[task.name] = labels
            all_predictions[task.name] = predictions
            if weights is not None:
                all_weights[task.name] = weights.squeeze()
    if required_inputs_list:
        for name in required_inputs_list:
            if name in model_out:
                all_required_inputs[name] = model_out[name].squeeze()
    return all_labels, all_predictions, all_weights, all_required_inputs


class RecMetricModule(nn.Module):
    """Owns the rec metrics, the throughput metric and state metrics.

    ``update(model_out)`` every step; ``compute()`` (rate-limited by ``compute_interval_steps`` /
    ``min_compute_interval``) returns one flat ``{key: value}`` dict with keys like ``ne-task|lifetime_ne``."""

    def __init__(self, batch_size: int, world_size: int, rec_tasks: Optional[List[RecTaskInfo]] = None, rec_metrics: Optional[RecMetricList] = None,
                 throughput_metric: Optional[ThroughputMetric] = None, state_metrics: Optional[Dict[str, StateMetric]] = None, compute_interval_steps: int = 100,
                 min_compute_interval: float = 0.0, max_compute_interval: float = float("inf")) -> None:
        super().__init__()
        self.rec_tasks = rec_tasks if rec_tasks else []
        self.rec_metrics = rec_metrics if rec_metrics else RecMetricList([])
        self.throughput_metric = throughput_metric
        self.state_metrics = state_metrics if state_metrics else {}
        self.trained_batches: int = 0
        self.batch_size = batch_size
        self.world_size = world_size
        self.oom_count = 0
        self.compute_count = 0
        self.compute_interval_steps = compute_interval_steps
        self.min_compute_interval = min_compute_interval
        self.max_compute_interval = max_compute_interval
        if self.min_compute_interval == 0.0 and self.max_compute_interval == float("inf"):
            self.min_compute_interval = -1.0
            self.max_compute_interval = -1.0
        else:
            if self.max_compute_interval <= 0.0:
                raise ValueError("Max compute interval should not be smaller than 0.0.")
            if self.min_compute_interval < 0.0:
                raise ValueError("Min compute interval should not be smaller than 0.0.")
        self.register_buffer("_compute_interval_steps", torch.zeros(1, dtype=torch.int32), persistent=False)
        self.last_compute_time = -1.0

    def _update_rec_metrics(self, model_out: Dict[str, torch.Tensor], **kwargs: Any) -> None:
        if self.rec_metrics and self.rec_tasks:
            labels, predictions, weights, required_inputs = parse_task_model_outputs(self.rec_tasks, model_out, self.rec_metrics.get_required_inputs())
            if required_inputs:
                kwargs["required_inputs"] = required_inputs
            self.rec_metrics.update(predictions=predictions, labels=labels, weights=weights if weights else None, **kwargs)

    def update(self, model_out: Dict[str, torch.Tensor], **kwargs: Any) -> None:
        with torch.no_grad():
            self._update_rec_metrics(model_out, **kwargs)
            if self.throughput_metric:
                self.throughput_metric.update()
            self.trained_batches += 1

    def _adjust_compute_interval(self) -> None:
        """Keep wall-clock between computes inside [min, max] by adapting the step interval (agreed across ranks)."""
        if self.last_compute_time > 0 and self.min_compute_interval >= 0:
            now = time.time()
            interval = (now - self.last_compute_time) / max(self.compute_interval_steps, 1)
            if not (self.max_compute_interval >= interval * self.compute_interval_steps >= self.min_compute_interval):
                per_step = max(interval, 1e-9)
                target = (self.min_compute_interval + min(self.max_compute_interval, self.min_compute_interval * 2 + 1)) / 2
                self._compute_interval_steps[0] = max(1, int(target / per_step))
                if dist.is_initialized():
                    dist.all_reduce(self._compute_interval_steps, op=dist.ReduceOp.MAX)
                self.compute_interval_steps = int(self._compute_interval_steps.item())
        self.last_compute_time = time.time()

    def should_compute(self) -> bool:
        return self.trained_batches % self.compute_interval_steps == 0

    def compute(self) -> Dict[str, MetricValue]:
        self.compute_count += 1
        ret: Dict[str, MetricValue] = {}
        with torch.no_grad():
            if self.rec_metrics:
                self._adjust_compute_interval()
                ret.update(self.rec_metrics.compute())
            if self.throughput_metric:
                ret.update(self.throughput_metric.compute())
            if self.state_metrics:
                for namespace, component in self.state_metrics.items():
                    ret.update({compose_customized_metric_key(namespace, metric_name): metric_value for metric_name, metric_value in component.get_metrics().items()})
        return ret

    def local_compute(self) -> Dict[str, MetricValue]:
        ret: Dict[str, MetricValue] = {}
        if self.rec_metrics:
            ret.update(self.rec_metrics.local_compute())
        if self.throughput_metric:
            ret.update(self.throughput_metric.compute())
        return ret

    def sync(self) -> None:
        pass

    def unsync(self) -> None:
        pass

    def reset(self) -> None:
        self.rec_metrics.reset()

    def get_required_inputs(self) -> Optional[List[str]]:
        return self.rec_metrics.get_required_inputs()

    def get_memory_usage(self) -> int:
        total = 0
        for m in self.rec_metrics.rec_metrics:
            total += sum(m.get_memory_usage().values())
        return total




def _generate_rec_metrics(metrics_config: MetricsConfig, world_size: int, my_rank: int, batch_size: int, process_group: Optional[dist.ProcessGroup] = None) -> RecMetricList:
    rec_metrics = []
    for metric_enum, metric_def in metrics_config.rec_metrics.items():
        kwargs: Dict[str, Any] = {}
        if metric_def and metric_def.arguments is not None:
            kwargs = metric_def.arguments
        rec_tasks: List[RecTaskInfo] = []
        if metric_def.rec_tasks and metric_def.rec_task_indices:
            raise ValueError("Only one of RecMetricDef.rec_tasks and RecMetricDef.rec_task_indices should be specified.")
        if metric_def.rec_tasks:
            rec_tasks = metric_def.rec_tasks
        elif metric_def.rec_task_indices:
            rec_tasks = [metrics_config.rec_tasks[idx] for idx in metric_def.rec_task_indices]
        else:
            raise ValueError("One of RecMetricDef.rec_tasks and RecMetricDef.rec_task_indices should be a non-empty list")
        rec_metrics.append(REC_METRICS_MAPPING[metric_enum](
            world_size=world_size, my_rank=my_rank, batch_size=batch_size, tasks=rec_tasks, compute_mode=metrics_config.rec_compute_mode,
            window_size=metric_def.window_size, fused_update_limit=metrics_config.fused_update_limit, compute_on_all_ranks=metrics_config.compute_on_all_ranks,
            should_validate_update=metrics_config.should_validate_update, process_group=process_group, **kwargs))
    return RecMetricList(rec_metrics)


STATE_METRICS_NAMESPACE_MAPPING: Dict[StateMetricEnum, MetricNamespace] = {
    StateMetricEnum.OPTIMIZERS: MetricNamespace.OPTIMIZERS,
    StateMetricEnum.MODEL_CONFIGURATOR: MetricNamespace.MODEL_CONFIGURATOR,
}


def _generate_state_metrics(metrics_config: MetricsConfig, state_metrics_mapping: Dict[StateMetricEnum, StateMetric]) -> Dict[str, StateMetric]:
    state_metrics: Dict[str, StateMetric] = {}
    for metric_enum in metrics_config.state_metrics:
        metric_namespace: Optional[MetricNamespace] = STATE_METRICS_NAMESPACE_MAPPING.get(metric_enum, None)
        if metric_namespace is None:
            raise ValueError(f"Unknown StateMetrics {metric_enum}")
        state_metrics[metric_namespace.value] = state_metrics_mapping[metric_enum]
    return state_metrics


def generate_metric_module(metric_class: Type[RecMetricModule], metrics_config: MetricsConfig, batch_size: int, world_size: int, my_rank: int,
                           state_metrics_mapping: Dict[StateMetricEnum, StateMetric], device: torch.device, process_group: Optional[dist.ProcessGroup] = None,
                           batching_metadata: Optional[Any] = None, batch_size_stages: Optional[List[Any]] = None, module_kwargs: Optional[Dict[str, Any]] = None) -> RecMetricModule:
    from .metrics_config import validate_batch_size_stages

    validate_batch_size_stages(batch_size_stages)
    rec_metrics = _generate_rec_metrics(metrics_config, world_size, my_rank, batch_size, process_group)
    throughput_metric = ThroughputMetric(batch_size=batch_size, world_size=world_size, window_seconds=metrics_config.throughput_metric.window_size) \
        if metrics_config.throughput_metric else None
    state_metrics = _generate_state_metrics(metrics_config, state_metrics_mapping)
    metrics = metric_class(batch_size=batch_size, world_size=world_size, rec_tasks=metrics_config.rec_tasks, rec_metrics=rec_metrics,
                           throughput_metric=throughput_metric, state_metrics=state_metrics, compute_interval_steps=metrics_config.compute_interval_steps,
                           min_compute_interval=metrics_config.min_compute_interval, max_compute_interval=metrics_config.max_compute_interval, **(module_kwargs or {}))
    metrics.to(device)
    return metrics


# ---- moved to ``cpu_offloaded_metric_module.py`` (their reference import path); still importable from here ----
_MOVED_TO_CPU_OFFLOADED_METRIC_MODULE = ('CPUOffloadedRecMetricModule',)


def __getattr__(name: str):
    if name in _MOVED_TO_CPU_OFFLOADED_METRIC_MODULE:
        from . import cpu_offloaded_metric_module as _m

        return getattr(_m, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
