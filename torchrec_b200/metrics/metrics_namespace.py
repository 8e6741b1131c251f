"""Metric naming (reference torchrec/metrics/metrics_namespace.py)."""
from enum import Enum
from typing import Optional


class StrValueMixin:
    def __str__(self) -> str:
        return self.value  # type: ignore[attr-defined]


class MetricNameBase(StrValueMixin, Enum):
    """Base of metric-name enums (projects add their own names by subclassing)."""


class MetricName(MetricNameBase):
    DEFAULT = ""
    NE = "ne"
    NE_POSITIVE = "ne_positive"
    LOG_LOSS = "logloss"
    THROUGHPUT = "throughput"
    TOTAL_EXAMPLES = "total_examples"
    ATTEMPT_EXAMPLES = "attempt_examples"
    BATCH_SIZE = "batch_size"
    CTR = "ctr"
    CALIBRATION = "calibration"
    MSE = "mse"
    MAE = "mae"
    RMSE = "rmse"
    AUC = "auc"
    AUPRC = "auprc"
    RAUC = "rauc"
    GROUPED_AUC = "grouped_auc"
    GROUPED_AUPRC = "grouped_auprc"
    RECALL_SESSION_LEVEL = "recall_session_level"
    PRECISION_SESSION_LEVEL = "precision_session_level"
    MULTICLASS_RECALL = "multiclass_recall"
    WEIGHTED_AVG = "weighted_avg"
    TOWER_QPS = "qps"
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
    NRMSE = "nrmse"
    AVERAGE = "average"
    SEGMENTED_NE = "segmented_ne"
    MULTI_LABEL_PRECISION = "multi_label_precision"
    RECALIBRATED_NE = "recalibrated_ne"
    RECALIBRATED_CALIBRATION = "recalibrated_calibration"
    SERVING_AE_LOSS = "serving_ae_loss"
    NUM_POSITIVE_SAMPLES = "num_positive_samples"
    SUM_WEIGHTS = "sum_weights"
    NUM_MISSING_LABELS = "num_missing_labels"
    WEIGHTED_SUM_PREDICTIONS = "weighted_sum_predictions"
    GAUC = "gauc"
    GAUC_NUM_SAMPLES = "gauc_num_samples"
    GROUPED_RAUC = "grouped_rauc"
    HINDSIGHT_TARGET_PRECISION = "hindsight_target_precision"
    HINDSIGHT_TARGET_RECALL = "hindsight_target_recall"
    LABEL_AVERAGE = "label_average"
    PREDICTION_AVERAGE = "prediction_average"
    DATA_LABEL_AVERAGE = "data_label_average"
    R_SQUARED = "r_squared"
    EFFECTIVE_RATE = "effective_rate"
    EFFECTIVE_SAMPLE_RATE = "effective_sample_rate"
    TOTAL_POSITIVE_EXAMPLES = "total_positive_examples"
    TOTAL_NEGATIVE_EXAMPLES = "total_negative_examples"


class MetricNamespaceBase(StrValueMixin, Enum):
    pass


class MetricNamespace(MetricNamespaceBase):
    DEFAULT = ""
    NE = "ne"
    NE_POSITIVE = "ne_positive"
    THROUGHPUT = "throughput"
    CTR = "ctr"
    CALIBRATION = "calibration"
    MSE = "mse"
    AUC = "auc"
    AUPRC = "auprc"
    RAUC = "rauc"
    MAE = "mae"
    ACCURACY = "accuracy"
    OPTIMIZERS = "optimizers"
    MODEL_CONFIGURATOR = "model_configurator"
    MULTICLASS_RECALL = "multiclass_recall"
    WEIGHTED_AVG = "weighted_avg"
    RECALL_SESSION_LEVEL = "recall_session_level"
    PRECISION_SESSION_LEVEL = "precision_session_level"
    TOWER_QPS = "qps"
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
    SEGMENTED_NE = "segmented_ne"
    GROUPED_AUC = "grouped_auc"
    MULTI_LABEL_PRECISION = "multi_label_precision"
    RECALIBRATED_NE = "recalibrated_ne"
    RECALIBRATED_CALIBRATION = "recalibrated_calibration"
    SERVING_AE_LOSS = "serving_ae_loss"
    NUM_POSITIVE_SAMPLES = "num_positive_samples"
    SUM_WEIGHTS = "sum_weights"
    NUM_MISSING_LABELS = "num_missing_labels"
    WEIGHTED_SUM_PREDICTIONS = "weighted_sum_predictions"
    GAUC = "gauc"


class MetricPrefix(StrValueMixin, Enum):
    DEFAULT = ""
    LIFETIME = "lifetime_"
    WINDOW = "window_"
    ATTEMPT = "attempt_"


def task_wildcard_metrics_pattern(namespace: MetricNamespaceBase, metric_name: MetricName, metric_prefix: MetricPrefix = MetricPrefix.DEFAULT) -> str:
    return f"{namespace}-.+\\|{metric_prefix}{metric_name}"


def compose_metric_namespace(namespace: MetricNamespaceBase, task_name: str) -> str:
    return f"{namespace}-{task_name}"


def compose_customized_metric_key(namespace: str, metric_name: str, description: Optional[str] = None) -> str:
    return f"{namespace}|{metric_name}{description or ''}"


def compose_metric_key(namespace: MetricNamespaceBase, task_name: str, metric_name: MetricName, metric_prefix: MetricPrefix = MetricPrefix.DEFAULT,
                       description: Optional[str] = None) -> str:
    """``<namespace>-<task>|<prefix><metric>`` e.g. ``ne-ctr_task|lifetime_ne``."""
    return compose_customized_metric_key(compose_metric_namespace(namespace, task_name), f"{metric_prefix}{metric_name}", description)
