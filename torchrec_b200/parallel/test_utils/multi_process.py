"""Multi-process harness (reference ``torchrec/distributed/test_utils/multi_process.py``: ``MultiProcessContext`` :31, ``MultiProcessTestBase`` :98,
``run_multi_process_func`` :232). The process / rendezvous logic is ``utils/multiprocess.py`` (spawn, 127.0.0.1 rendezvous, gloo on CPU / nccl on GPUs)."""
from __future__ import annotations

import unittest
from typing import Any, Callable, Dict, List, Optional

from ...utils.multiprocess import MultiProcessContext, run_multi_process  # noqa: F401


class MultiProcessTestBase(unittest.TestCase):
    """``self._run_multi_process_test(callable=fn, world_size=2, **kwargs)`` runs ``fn(rank, world_size, **kwargs)`` in ``world_size`` processes and fails the
    test if any of them raises."""

    backend: Optional[str] = None

    def _run_multi_process_test(self, *, callable: Callable[..., None], world_size: int = 2, local_size: Optional[int] = None, **kwargs: Any) -> None:
        run_multi_process(_RankWorldAdapter(callable), world_size=world_size, backend=self.backend or _default_backend(world_size), local_size=local_size, **kwargs)

    def _run_multi_process_test_per_rank(self, *, callable: Callable[..., None], world_size: int, kwargs_per_rank: List[Dict[str, Any]]) -> None:
        run_multi_process(_RankWorldAdapter(callable, kwargs_per_rank), world_size=world_size, backend=self.backend or _default_backend(world_size))


def _default_backend(world_size: int) -> str:
    import torch

    return "nccl" if torch.cuda.is_available() and torch.cuda.device_count() >= world_size else "gloo"


class _RankWorldAdapter:
    """picklable ``fn(ctx, **kw)`` -> ``callable(rank, world_size, **kw)``"""

    def __init__(self, fn: Callable[..., None], kwargs_per_rank: Optional[List[Dict[str, Any]]] = None) -> None:
        self.fn, self.kwargs_per_rank = fn, kwargs_per_rank

    def __call__(self, ctx: MultiProcessContext, **kwargs: Any) -> None:
        if self.kwargs_per_rank is not None:
            kwargs = {**kwargs, **self.kwargs_per_rank[ctx.rank]}
        self.fn(ctx.rank, ctx.world_size, **kwargs)


def run_multi_process_func(func: Callable[..., None], world_size: int = 2, multiprocessing_method: str = "spawn", backend: Optional[str] = None,
                           local_size: Optional[int] = None, **kwargs: Any) -> None:
    """``func(rank, world_size, **kwargs)`` in every process (the benchmarks' entry point)."""
    run_multi_process(_RankWorldAdapter(func), world_size=world_size, backend=backend or _default_backend(world_size), local_size=local_size, **kwargs)
