"""Gradient clipping optimizer wrapper (reference torchrec/optim/clipping.py:32): clip by norm or value,
sharded-parameter aware (norms of ShardedTensor / local shards are summed across ranks)."""
from __future__ import annotations

from enum import Enum, unique
from typing import Any, Dict, List, Optional, Union

import torch
import torch.distributed as dist
from torch.distributed._shard.sharded_tensor import ShardedTensor

from .keyed import KeyedOptimizer, OptimizerWrapper


@unique
class GradientClipping(Enum):
    NORM = "norm"
    VALUE = "value"
    NONE = "none"


class GradientClippingOptimizer(OptimizerWrapper):
    """Clips gradients before the wrapped optimizer steps.

    ``enable_global_grad_clip``: the norm is computed over *all* parameters of the job — sharded parameters
    contribute their local shard norms which are all-reduced over ``sharded_pg`` (replicated ones are counted
    once)."""

    def __init__(self, optimizer: KeyedOptimizer, clipping: GradientClipping = GradientClipping.NONE, max_gradient: float = 0.1,
                 norm_type: Union[float, str] = 2.0, enable_global_grad_clip: bool = False,
                 param_to_pgs: Optional[Dict[torch.nn.Parameter, List[dist.ProcessGroup]]] = None) -> None:
        super().__init__(optimizer)
        self._clipping = clipping
        self._max_gradient = max_gradient
        self._norm_type = float(norm_type)
        self._check_meta: bool = True
        self._enable_global_grad_clip = enable_global_grad_clip
        self._step_num = 0
        self._params: List[torch.Tensor] = []
        self._sharded_params: List[torch.Tensor] = []
        self._replicate_params: List[torch.Tensor] = []
        for param_group in self.param_groups:
            self._params += list(param_group["params"])
        for p in self._params:
            (self._sharded_params if isinstance(p, ShardedTensor) or getattr(p, "_is_sharded", False) else self._replicate_params).append(p)

    def step(self, closure: Any = None) -> None:
        if self._check_meta:
            if any(getattr(param, "is_meta", False) for param in self._params):
                return
            self._check_meta = False
        if self._clipping == GradientClipping.NORM:
            if self._enable_global_grad_clip:
                self.clip_grad_norm_()
            else:
                grads = [p for p in self._replicate_params if p.grad is not None]
                if grads:
                    torch.nn.utils.clip_grad_norm_(grads, self._max_gradient, norm_type=self._norm_type)
        elif self._clipping == GradientClipping.VALUE:
            torch.nn.utils.clip_grad_value_([p for p in self._replicate_params if p.grad is not None], self._max_gradient)
        super().step(closure)
        self._step_num += 1

    @torch.no_grad()
    def clip_grad_norm_(self) -> Optional[torch.Tensor]:
        """Global norm over sharded + replicated parameters; returns the total norm."""
        p = self._norm_type

        def local_pow_sum(params: List[torch.Tensor]) -> Optional[torch.Tensor]:
            acc = None
            for t in params:
                g = t.grad
                if g is None:
                    continue
                if isinstance(g, ShardedTensor):
                    vals = [s.tensor for s in g.local_shards()]
                else:
                    vals = [g]
                for v in vals:
                    n = torch.linalg.vector_norm(v.float(), p)
                    n = n if p == float("inf") else n**p
                    acc = n if acc is None else (torch.maximum(acc, n) if p == float("inf") else acc + n)
            return acc

        sharded = local_pow_sum(self._sharded_params)
        replicated = local_pow_sum(self._replicate_params)
        if sharded is not None and dist.is_initialized():
            dist.all_reduce(sharded, op=dist.ReduceOp.MAX if p == float("inf") else dist.ReduceOp.SUM)
        parts = [x for x in (sharded, replicated) if x is not None]
        if not parts:
            return None
        total = torch.stack(parts).max() if p == float("inf") else torch.stack(parts).sum() ** (1.0 / p)
        clip_coef = torch.clamp(self._max_gradient / (total + 1e-6), max=1.0)
        for t in self._sharded_params + self._replicate_params:
            if t.grad is None:
                continue
            if isinstance(t.grad, ShardedTensor):
                for s in t.grad.local_shards():
                    s.tensor.mul_(clip_coef)
            else:
                t.grad.mul_(clip_coef.to(t.grad.dtype))
        return total
