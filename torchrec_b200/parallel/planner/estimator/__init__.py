"""Annotation-driven perf estimator (reference ``torchrec/distributed/planner/estimator/``)."""
from .annotations import *  # noqa: F401,F403
from .config import EmbeddingPerfEstimatorConfig, GB200PerfConfig  # noqa: F401
from .estimator import (  # noqa: F401
    ColumnWiseEvaluator,
    DataParallelEvaluator,
    EmbeddingPerfEstimator,
    EmbeddingPerfEstimatorFactory,
    EmbeddingPerfEstimatorV2,
    EmbeddingShardingPerfEvaluator,
    GridShardEvaluator,
    RowWiseEvaluator,
    TableColumnWiseEvaluator,
    TableRowWiseEvaluator,
    TableWiseEvaluator,
    compute_block_usage_penalty,
    get_embedding_perf_sharding_evaluator,
)
from .types import EstimatorPerfCoefficients, HardwarePerfConfig, PerfCoefficient, PerfCoefficientConfig, PrefetchCoefficients, ShardPerfContext  # noqa: F401
