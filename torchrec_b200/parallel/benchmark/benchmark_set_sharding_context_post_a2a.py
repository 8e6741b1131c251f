"""Host-side micro-benchmark of ``_set_sharding_context_post_a2a`` (the bookkeeping after the KJT all-to-all that records the per-rank,
per-feature batch sizes of variable-batch inputs): the transposing list comprehension against a variant that re-reads
``stride_per_key_per_rank()`` inside the loops. Pure python cost per training step - it matters when there are hundreds of features.
Reference: ``distributed/benchmark/benchmark_set_sharding_context_post_a2a.py``.

    python -m torchrec_b200.distributed.benchmark.benchmark_set_sharding_context_post_a2a --num_list 0 --num_keys 0"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Any, Callable, Dict, List

import torch

from ...benchmarks.base import cmd_conf
from ...sparse.jagged_tensor import KeyedJaggedTensor
from ..embedding_sharding import _set_sharding_context_post_a2a


class _Ctx:
    def __init__(self, n: int) -> None:
        self.sharding_contexts = [type("S", (), {"batch_size_per_rank_per_feature": []})() for _ in range(n)]


def _set_sharding_context_post_a2a_previous(kjts: List[KeyedJaggedTensor], ctx: Any) -> None:
    """The straightforward form: every element access calls ``stride_per_key_per_rank()`` again."""
    for kjt, sctx in zip(kjts, getattr(ctx, "sharding_contexts", [])):
        if hasattr(sctx, "batch_size_per_rank_per_feature") and kjt.variable_stride_per_key() and kjt.stride_per_key_per_rank():
            sctx.batch_size_per_rank_per_feature = [[kjt.stride_per_key_per_rank()[i][j] for i in range(len(kjt.stride_per_key_per_rank()))]
                                                    for j in range(len(kjt.stride_per_key_per_rank()[0]))]


def op_bench(num_list: int, num_keys: int, func: Callable[..., None], world_size: int = 8, repeats: int = 5) -> Dict[str, float]:
    kjts = [KeyedJaggedTensor(keys=[f"k{i}" for i in range(num_keys)], values=torch.zeros(num_keys * world_size, dtype=torch.int64),
                              lengths=torch.ones(num_keys * world_size, dtype=torch.int64), stride_per_key_per_rank=[[1] * world_size for _ in range(num_keys)])
            for _ in range(num_list)]
    best = float("inf")
    for _ in range(repeats):
        ctx = _Ctx(num_list)
        t0 = time.perf_counter()
        func(kjts, ctx)
        best = min(best, time.perf_counter() - t0)
    assert ctx.sharding_contexts[0].batch_size_per_rank_per_feature == [[1] * num_keys for _ in range(world_size)]
    res = {"num_list": num_list, "num_keys": num_keys, "ms": best * 1e3}
    print(f"{func.__name__:45} lists={num_list:6} keys={num_keys:4}  {best * 1e3:9.3f} ms")
    return res


@dataclass
class PostA2AConfig:
    num_list: int = 100
    num_keys: int = 100


@cmd_conf
def main(cfg: PostA2AConfig) -> List[Dict[str, float]]:
    grid = [(nl, nk) for nl in (100, 1000, 10000) for nk in (10, 100)] if cfg.num_list == 0 and cfg.num_keys == 0 else [(cfg.num_list, cfg.num_keys)]
    out = []
    for nl, nk in grid:
        out.append(op_bench(nl, nk, _set_sharding_context_post_a2a_previous))
        out.append(op_bench(nl, nk, _set_sharding_context_post_a2a))
    return out


if __name__ == "__main__":
    main()
