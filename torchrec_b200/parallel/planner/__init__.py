"""Automatic sharding planner (reference torchrec/distributed/planner)."""
from .enumerators import EmbeddingEnumerator, EmbeddingPerfEstimator, EmbeddingStorageEstimator  # noqa: F401
from .partitioners import GreedyPerfPartitioner, MemoryBalancedPartitioner  # noqa: F401
from .perf_models import NoopCriticalPathPerfModel, NoopPerfModel  # noqa: F401
from .planners import EmbeddingShardingPlanner, HeteroEmbeddingShardingPlanner, to_sharding_plan  # noqa: F401
from .proposers import (  # noqa: F401
    DynamicProgrammingProposer,
    EmbeddingOffloadScaleupProposer,
    GreedyProposer,
    GridSearchProposer,
    UniformProposer,
)
from .stats import EmbeddingStats, NoopEmbeddingStats  # noqa: F401
from .storage_reservations import (  # noqa: F401
    FixedPercentageStorageReservation,
    HeuristicalStorageReservation,
    InferenceStorageReservation,
)
from .types import (  # noqa: F401
    ParameterConstraints,
    Perf,
    PlannerError,
    PlannerErrorType,
    Shard,
    ShardingOption,
    Storage,
    Topology,
)
