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


def create_on_rank_and_share_result(pg: dist.ProcessGroup, rank: int, creator: Callable[..., Any], extractor: Optional[Callable[[Any], List[Optional[torch.Tensor]]]] = None,
                                    constructor: Optional[Callable[[List[Optional[torch.Tensor]]], Any]] = None, *args: Any, **kwargs: Any) -> Any:
    """Run ``creator(*args, **kwargs)`` on ONE rank of a host-local group and hand the result to the other ranks through POSIX shared memory
    (no copy): ``extractor(result)`` lists the CPU tensors of the result, they are moved to shared memory and broadcast as handles,
    ``constructor(tensors)`` rebuilds the result on every rank. Without extractor / constructor the result must be a single CPU tensor.
    ``pg`` should be an intra-node group (``comm.intra_and_cross_node_pg``): shared memory is per host."""
    if pg.rank() == rank:
        result = creator(*args, **kwargs)
        tensors = extractor(result) if extractor is not None else [result]
        payload: List[Any] = [[t.share_memory_() if isinstance(t, torch.Tensor) else t for t in tensors]]
    else:
        payload = [None]
    dist.broadcast_object_list(payload, dist.get_global_rank(pg, rank), group=pg)
    tensors = payload[0]
    return constructor(tensors) if constructor is not None else tensors[0]
