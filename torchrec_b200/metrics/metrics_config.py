"""Metric configuration dataclasses (reference torchrec/metrics/metrics_config.py:21-250)."""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, List, Optional

from .rec_metric import RecComputeMode, RecTaskInfo


class RecMetricEnumBase(Enum):
    pass


class RecMetricEnum(RecMetricEnumBase):
    NE = "ne"
    NE_POSITIVE = "ne_positive"
    SEGMENTED_NE = "segmented_ne"
    LOG_LOSS = "log_loss"
    CTR = "ctr"
    AUC = "auc"
    AUPRC = "auprc"
    RAUC = "rauc"
    CALIBRATION = "calibration"
    MSE = "mse"
    MAE = "mae"
    MULTICLASS_RECALL = "multiclass_recall"
    RECALL_SESSION_LEVEL = "recall_session_level"
    PRECISION_SESSION_LEVEL = "precision_session_level"
    WEIGHTED_AVG = "weighted_avg"
    TOWER_QPS = "tower_qps"
    ACCURACY = "accuracy"
    NDCG = "ndcg"
    XAUC = "xauc"
    SCALAR = "scalar"
    PRECISION = "precision"
    RECALL = "recall"
    SERVING_NE = "serving_ne"
    SERVING_CALIBRATION = "serving_calibration"
    OUTPUT = "output"
    TENSOR_WEIGHTED_AVG = "tensor_weighted_avg"
    CALI_FREE_NE = "cali_free_ne"
    UNWEIGHTED_NE = "unweighted_ne"
    HINDSIGHT_TARGET_PR = "hindsight_target_pr"
    NMSE = "nmse"
    AVERAGE = "average"
    GAUC = "gauc"
    MULTI_LABEL_PRECISION = "multi_label_precision"
    RECALIBRATED_NE = "recalibrated_ne"
    RECALIBRATED_CALIBRATION = "recalibrated_calibration"
    SERVING_AE_LOSS = "serving_ae_loss"
    NUM_POSITIVE_SAMPLES = "num_positive_samples"
    SUM_WEIGHTS = "sum_weights"
    NUM_MISSING_LABELS = "num_missing_labels"
    WEIGHTED_SUM_PREDICTIONS = "weighted_sum_predictions"


@dataclass(unsafe_hash=True, eq=True)
class SessionMetricDef:
    session_var_name: str
    top_threshold: Optional[int] = None
    run_ranking_of_labels: bool = False


@dataclass
class RecMetricDef:
    rec_tasks: List[RecTaskInfo] = field(default_factory=list)
    rec_task_indices: List[int] = field(default_factory=list)
    window_size: int = 10_000_000
    arguments: Optional[Dict[str, Any]] = None


class StateMetricEnum(Enum):
    OPTIMIZERS = "optimizers"
    MODEL_CONFIGURATOR = "model_configurator"


@dataclass
class ThroughputDef:
    window_size: int = 100


@dataclass
class BatchSizeStage:
    batch_size: int
    max_iters: Optional[int] = None


@dataclass
class MetricsConfig:
    rec_tasks: List[RecTaskInfo] = field(default_factory=list)
    rec_metrics: Dict[RecMetricEnumBase, RecMetricDef] = field(default_factory=dict)
    throughput_metric: Optional[ThroughputDef] = None
    rec_compute_mode: RecComputeMode = RecComputeMode.UNFUSED_TASKS_COMPUTATION
    fused_update_limit: int = 0
    state_metrics: List[StateMetricEnum] = field(default_factory=list)
    compute_interval_steps: int = 100
    min_compute_interval: float = 0.0
    max_compute_interval: float = float("inf")
    compute_on_all_ranks: bool = False
    should_validate_update: bool = False


DefaultTaskInfo = RecTaskInfo(name="DefaultTask", label_name="label", prediction_name="prediction", weight_name="weight")

DefaultMetricsConfig = MetricsConfig(
    rec_tasks=[DefaultTaskInfo],
    rec_metrics={RecMetricEnum.NE: RecMetricDef(rec_tasks=[DefaultTaskInfo], window_size=10_000_000)},
    throughput_metric=ThroughputDef(),
    state_metrics=[],
)

EmptyMetricsConfig = MetricsConfig(rec_tasks=[], rec_metrics={}, throughput_metric=None, state_metrics=[])


def validate_batch_size_stages(batch_size_stages: Optional[List[BatchSizeStage]]) -> None:
    """A batch-size schedule is a list of stages; only the last one may (and must) be open ended (``max_iters=None``)."""
    if not batch_size_stages:
        return
    if batch_size_stages[-1].max_iters is not None:
        raise ValueError(f"Batch size stages last stage should have max_iters = None, but get {batch_size_stages[-1].max_iters}")
    if any(stage.max_iters is None for stage in batch_size_stages[:-1]):
        raise ValueError("Batch size stages should have max_iters set for every stage but the last")
