"""Small collective helpers (reference torchrec/distributed/collective_utils.py:40,77)."""
from typing import Any, Callable, List, Optional, TypeVar

import torch
import torch.distributed as dist

T = TypeVar("T")


def is_leader(pg: Optional[dist.ProcessGroup], leader_rank: int = 0) -> bool:
    if pg is None:
        return leader_rank == 0
    return pg.rank() == leader_rank


def invoke_on_rank_and_broadcast_result(pg: dist.ProcessGroup, rank: int, func: Callable[..., T], *args: Any, **kwargs: Any) -> T:
    """Run ``func`` on one rank and broadcast its (picklable) result to every rank of ``pg``."""
    if pg.rank() == rank:
        res = func(*args, **kwargs)
        object_list = [res]
    else:
        object_list = [None]
    if pg.size() > 1:
        src = dist.get_global_rank(pg, rank) if hasattr(dist, "get_global_rank") else rank
        dist.broadcast_object_list(object_list, src, group=pg)
    return object_list[0]


def run_on_leader(pg: dist.ProcessGroup, rank: int):
    def decorator(func: Callable[..., T]) -> T:
        def wrapped(*args: Any, **kwargs: Any) -> T:
            return invoke_on_rank_and_broadcast_result(pg, rank, func, *args, **kwargs)

        return wrapped

    return decorator


def create_on_rank_and_share_result(pg: dist.ProcessGroup, rank: int, tensor_builder: Callable[[], torch.Tensor]) -> torch.Tensor:
    """Build a CPU tensor on one rank and share it with local peers through POSIX shared memory."""
    if pg.rank() == rank:
        t = tensor_builder().share_memory_()
        payload: List[Any] = [t]
    else:
        payload = [None]
    dist.broadcast_object_list(payload, dist.get_global_rank(pg, rank), group=pg)
    return payload[0]
