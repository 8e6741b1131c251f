from .metric_module import CPUOffloadedRecMetricModule, RecMetricModule, StateMetric, generate_metric_module  # noqa: F401
from .metrics_config import DefaultMetricsConfig, DefaultTaskInfo, MetricsConfig, RecMetricDef, RecMetricEnum, StateMetricEnum, ThroughputDef  # noqa: F401
from .metrics_impl import *  # noqa: F401,F403
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix, compose_metric_key  # noqa: F401
from .rec_metric import RecComputeMode, RecMetric, RecMetricComputation, RecMetricList, RecTaskInfo  # noqa: F401
from .throughput import ThroughputMetric  # noqa: F401
from .cpu_comms_metric_module import CPUCommsRecMetricModule  # noqa: F401,E402
from .deferrable_metrics import DeferrableMetrics  # noqa: F401,E402
from .metric_job_types import MetricComputeJob, MetricUpdateJob, SynchronizationMarker  # noqa: F401,E402
from .metric_state_snapshot import MetricStateSnapshot  # noqa: F401,E402
from .noop_metric_module import NoOpMetricModule  # noqa: F401,E402
