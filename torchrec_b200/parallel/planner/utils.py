"""Planner helpers."""
import math
import operator
from functools import reduce
from typing import Any, Dict, Iterable, List, Optional, Tuple, Type, Union

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


def mb_to_bytes(mb: float) -> int:
    return int(mb * 1024 * 1024)


class LuusJaakolaSearch:
    """Clamped Luus-Jaakola random search for the minimum of a 1-D function on [A, B] (the proposers tune a memory budget with it):
    ``next(f(previous point))`` returns the next point to evaluate (the first call ignores its argument), None after
    ``max_iterations``; a better point shrinks the sampling radius by 5 %. Samples never leave [A, B]; a boundary whose cost is already
    known (``left_cost``, ``shrink_right``) is not sampled again."""

    def __init__(self, A: float, B: float, max_iterations: int, seed: int = 42, left_cost: Optional[float] = None) -> None:
        import math

        self.left, self.right = A, B
        self.iteration = -1
        self.max_iterations = max_iterations
        self.gen = torch.Generator()
        self.gen.manual_seed(seed)
        self.x: float = self.uniform(self.left, self.right)
        self.fx: float = 0.0
        self.y: float = math.nan
        self.fleft: Optional[float] = left_cost
        self.fright: Optional[float] = None
        self.d: float = self.right - self.left

    def shrink_right(self, B: float) -> None:
        """Everything at and beyond ``B`` is infeasible (infinite cost)."""
        import math

        self.right = B
        self.fright = math.inf
        self.d = self.right - self.left
        self.x = self.clamp(self.x)

    def clamp(self, x: float) -> float:
        return self.left if x < self.left else self.right if x > self.right else x

    def uniform(self, A: float, B: float) -> float:
        return A + (B - A) * torch.rand(1, generator=self.gen, device="cpu").item()

    def next(self, fy: float) -> Optional[float]:
        self.iteration += 1
        if self.iteration == 0:
            return self.x
        if self.iteration == 1:
            self.fx = fy
        elif self.iteration == self.max_iterations:
            return None
        elif fy <= self.fx:
            self.x, self.fx = self.y, fy
            self.d = 0.95 * self.d
        if self.y == self.left:
            self.fleft = fy
        elif self.y == self.right:
            self.fright = fy
        if self.d <= 0 or (self.left == self.right):
            return None
        while True:
            y = self.clamp(self.x + self.uniform(-self.d, self.d))
            if (y == self.left and self.fleft is not None) or (y == self.right and self.fright is not None):
                continue
            self.y = y
            return y

    def best(self) -> Tuple[float, float]:
        return self.x, self.fx


# ---- picklable sharder snapshots for the estimators --------------------------------------------------------------------------------------------------
def build_sharder_data(sharder: Any) -> "SharderData":
    """What the estimators read from a sharder (fused params, byte sizes of quantized collectives, which storage formula applies)
    without keeping the sharder object alive."""
    from ..embedding_types import BaseEmbeddingSharder, BaseQuantEmbeddingSharder
    from ..types import CommOp
    from .types import SharderData, StorageUsageType

    fused = dict(sharder.fused_params) if getattr(sharder, "fused_params", None) else {}
    sizes: Dict[str, Tuple[float, float]] = {}
    registry = getattr(sharder, "qcomm_codecs_registry", None)
    if registry is not None:
        for op in (CommOp.POOLED_EMBEDDINGS_ALL_TO_ALL, CommOp.SEQUENCE_EMBEDDINGS_ALL_TO_ALL, CommOp.POOLED_EMBEDDINGS_REDUCE_SCATTER):
            if op.name in registry:
                codecs = registry[op.name]
                sizes[op.name] = (torch.tensor([], dtype=codecs.forward.quantized_dtype).element_size(), torch.tensor([], dtype=codecs.backward.quantized_dtype).element_size())
    usage = StorageUsageType.BASE_QUANT if isinstance(sharder, BaseQuantEmbeddingSharder) else StorageUsageType.BASE if isinstance(sharder, BaseEmbeddingSharder) \
        else StorageUsageType.DEFAULT
    return SharderData(fused_params=fused, qcomm_dtype_sizes=sizes, storage_usage_type=usage)


def build_sharder_data_map(sharder_map: Dict[str, Any]) -> Dict[str, "SharderData"]:
    return {key: build_sharder_data(sharder) for key, sharder in sharder_map.items()}


def is_prefetch_pipelined(sharding_option: ShardingOption, sharder_data: Any) -> bool:
    """Cache prefetch overlapped with the previous step: the option's cache params say so, or the sharder's fused params."""
    flag = sharding_option.cache_params.prefetch_pipeline if sharding_option.cache_params else None
    if not flag:
        fused = sharder_data.fused_params if hasattr(sharder_data, "fused_params") else (sharder_data or {})
        flag = (fused or {}).get("prefetch_pipeline", False)
    return bool(flag)


def extract_comm_data_type_size(sharding_option: ShardingOption, sharder_data: Any) -> Tuple[float, float, float, float]:
    """Bytes per element on the wire: (forward all-to-all, backward all-to-all, forward reduce-scatter, backward reduce-scatter) - the
    table's element size unless the sharder quantizes that collective."""
    from ..types import CommOp

    size = sharding_option.tensor.element_size()
    fwd_a2a = bwd_a2a = fwd_sr = bwd_sr = size
    qcomm = getattr(sharder_data, "qcomm_dtype_sizes", {}) or {}
    a2a = CommOp.POOLED_EMBEDDINGS_ALL_TO_ALL.name if sharding_option.is_pooled else CommOp.SEQUENCE_EMBEDDINGS_ALL_TO_ALL.name
    if a2a in qcomm:
        fwd_a2a, bwd_a2a = qcomm[a2a]
    if sharding_option.is_pooled and CommOp.POOLED_EMBEDDINGS_REDUCE_SCATTER.name in qcomm:
        fwd_sr, bwd_sr = qcomm[CommOp.POOLED_EMBEDDINGS_REDUCE_SCATTER.name]
    return fwd_a2a, bwd_a2a, fwd_sr, bwd_sr


def get_num_poolings(constraints: Optional[Dict[str, Any]], so: ShardingOption) -> List[float]:
    """Poolings per feature of an option: its own ``num_poolings`` when it matches the inputs, else the constraint's, else 1.0 each."""
    own = getattr(so, "num_poolings", None)
    if own is not None and len(own) == len(so.input_lengths):
        return list(own)
    if constraints and constraints.get(so.name) is not None and constraints[so.name].num_poolings:
        return list(constraints[so.name].num_poolings)
    from .constants import NUM_POOLINGS

    return [NUM_POOLINGS] * len(so.input_lengths)
