"""Plan rating (reference planner/perf_models.py): the slowest device decides the step time."""
from typing import List, cast

from .types import Perf, PerfModel, ShardingOption, Topology


class NoopPerfModel(PerfModel):
    def __init__(self, topology: Topology) -> None:
        self._topology = topology

    def rate(self, plan: List[ShardingOption]) -> float:
        perfs = [0.0] * self._topology.world_size
        for sharding_option in plan:
            for shard in sharding_option.shards:
                assert shard.rank is not None
                perfs[shard.rank] += cast(Perf, shard.perf).total
        return max(perfs)


class NoopCriticalPathPerfModel(PerfModel):
    """Sums the slowest device of every phase (comms / compute happen in lock-step across ranks)."""

    def __init__(self, topology: Topology, comms_group_keys=None, comp_group_keys=None) -> None:
        self._topology = topology

    def rate(self, plan: List[ShardingOption]) -> float:
        W = self._topology.world_size
        fc, fcm, bc, bcm = [0.0] * W, [0.0] * W, [0.0] * W, [0.0] * W
        for so in plan:
            for shard in so.shards:
                p = cast(Perf, shard.perf)
                r = shard.rank
                assert r is not None
                fc[r] += p.fwd_compute
                fcm[r] += p.fwd_comms + p.input_dist_comms
                bc[r] += p.bwd_compute
                bcm[r] += p.bwd_comms
        return max(fc) + max(fcm) + max(bc) + max(bcm)


class NoopStorageModel(PerfModel):
    """Rates a plan by the largest HBM use of any rank (lower = more evenly spread memory) without running anything."""

    def __init__(self, topology: Topology) -> None:
        self._topology = topology

    def rate(self, plan: List[ShardingOption]) -> float:
        hbm = [0] * self._topology.world_size
        for so in plan:
            for shard in so.shards:
                hbm[shard.rank] += shard.storage.hbm
        return max(hbm)
