"""Planner helpers."""
import math
import operator
from functools import reduce
from typing import Any, Dict, Iterable, List, Optional, Type, Union

import torch

from .types import Perf, ShardingOption, Storage


def sharder_name(t: Type[Any]) -> str:
    return t.__module__ + "." + t.__name__


def bytes_to_gb(num_bytes: int) -> float:
    return float(num_bytes / (1024 * 1024 * 1024))


def bytes_to_mb(num_bytes: Union[float, int]) -> float:
    return float(num_bytes / (1024 * 1024))


def gb_to_bytes(gb: float) -> int:
    return int(gb * 1024 * 1024 * 1024)


def prod(iterable: Iterable[int]) -> int:
    return reduce(operator.mul, iterable, 1)


def placement(compute_device: str, rank: int, local_size: int) -> str:
    from ..sharding_plan import placement as _p

    return _p(compute_device, rank, local_size)


def storage_repr_in_gb(storage: Optional[Storage]) -> str:
    if storage is None:
        return ""
    return f"Storage(hbm = {round(bytes_to_gb(storage.hbm), 3)} GB, ddr = {round(bytes_to_gb(storage.ddr), 3)} GB)"


def reset_shard_rank(proposal: List[ShardingOption]) -> None:
    for sharding_option in proposal:
        for shard in sharding_option.shards:
            shard.rank = None


def _find_imbalance_tables(sharding_options: List[ShardingOption], target_imbalance: str = "perf") -> List[ShardingOption]:
    rank_to_target_stats: Dict[int, float] = {}
    for so in sharding_options:
        for shard in so.shards:
            if shard.rank is None:
                continue
            v = shard.perf.total if target_imbalance == "perf" else shard.storage.hbm  # type: ignore[union-attr]
            rank_to_target_stats[shard.rank] = rank_to_target_stats.get(shard.rank, 0.0) + v
    if not rank_to_target_stats:
        return []
    max_rank = max(rank_to_target_stats, key=rank_to_target_stats.get)  # type: ignore[arg-type]
    return [so for so in sharding_options if any(shard.rank == max_rank for shard in so.shards)]


class BinarySearchPredicate:
    """Binary search over ints driven by a boolean predicate (used by the scale-up proposer)."""

    def __init__(self, A: int, B: int, tolerance: int) -> None:
        self.left = A
        self.right = B
        self.tolerance = tolerance
        self.first = True

    def next(self, prior_result: bool) -> Optional[int]:
        if self.right - self.left < self.tolerance:
            return None
        mid = self._mid()
        if self.first:
            self.first = False
            return mid
        if prior_result:
            self.left = mid + 1
        else:
            self.right = mid - 1
        if self.right - self.left < self.tolerance:
            return None
        return self._mid()

    def _mid(self) -> int:
        return self.left + ((self.right - self.left) // 2)
