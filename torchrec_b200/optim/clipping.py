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
            (self._sharded_params if self._is_sharded(p) else self._replicate_params).append(p)
        self._param_to_pgs = param_to_pgs or {}
        self._reduce_device: Optional[torch.device] = None
        # does ANY rank own sharded parameters under this optimizer? (decided once, collectively, so that clip_grad_norm_ issues the
        # same all-reduces everywhere; without a process group the local answer is the global one)
        self._has_sharded_anywhere = bool(self._sharded_params)
        if enable_global_grad_clip and dist.is_initialized() and dist.get_world_size() > 1:
            flags: List[Any] = [None] * dist.get_world_size()
            dist.all_gather_object(flags, bool(self._sharded_params))
            self._has_sharded_anywhere = any(flags)
            if dist.get_backend() == "nccl" and torch.cuda.is_available():
                self._reduce_device = torch.device("cuda", torch.cuda.current_device())

    @staticmethod
    def _is_sharded(p: Any) -> bool:
        if isinstance(p, ShardedTensor) or getattr(p, "_is_sharded", False):
            return True
        placements = getattr(p, "placements", None)  # DTensor: sharded unless every placement replicates
        return placements is not None and any(type(pl).__name__ != "Replicate" for pl in placements)

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

    def _local_grads(self, t: torch.Tensor) -> List[torch.Tensor]:
        g = t.grad
        if g is None:
            return []
        if isinstance(g, ShardedTensor):
            return [s.tensor for s in g.local_shards()]
        to_local = getattr(g, "to_local", None)  # DTensor gradient: this rank's shard(s)
        if to_local is not None:
            loc = to_local()
            return list(loc.local_shards()) if hasattr(loc, "local_shards") else [loc]
        return [g]

    def _groups_of_sharded_params(self) -> List[Optional[dist.ProcessGroup]]:
        """The process groups the sharded norms have to be summed over: every distinct group list of ``param_to_pgs`` (2D parallel:
        sharding group, then replica group is NOT reduced over - replicas hold the same shard), the default group otherwise. Computed
        from the parameter set, not from which gradients happen to exist, so every rank issues the same collectives."""
        if not self._param_to_pgs:
            return [None]
        seen: List[Optional[dist.ProcessGroup]] = []
        for t in self._sharded_params:
            for pg in self._param_to_pgs.get(t, [None]):
                if not any(pg is s for s in seen):
                    seen.append(pg)
        return seen or [None]

    @torch.no_grad()
    def clip_grad_norm_(self) -> Optional[torch.Tensor]:
        """Global norm over sharded + replicated parameters; returns the total norm. Collective: EVERY rank of the job takes part in
        the reduction of the sharded part whenever the optimizer owns sharded parameters anywhere (a rank without local shards or without
        gradients this step contributes zero) - the decision must not depend on rank-local state or the all-reduce deadlocks
        (reference optim/clipping.py:204-311)."""
        p = self._norm_type
        inf = p == float("inf")

        def local_pow_sum(params: List[torch.Tensor]) -> Optional[torch.Tensor]:
            acc = None
            for t in params:
                for v in self._local_grads(t):
                    n = torch.linalg.vector_norm(v.float(), p)
                    n = n if inf else n**p
                    acc = n if acc is None else (torch.maximum(acc, n) if inf else acc + n)
            return acc

        sharded = local_pow_sum(self._sharded_params)
        replicated = local_pow_sum(self._replicate_params)
        if dist.is_initialized() and (self._has_sharded_anywhere or sharded is not None):
            if sharded is None:
                ref = next((t for t in self._params if isinstance(t, torch.Tensor) and not getattr(t, "is_meta", False)), None)
                dev = self._reduce_device or (ref.device if ref is not None and not isinstance(ref, ShardedTensor) else torch.device("cpu"))
                sharded = torch.zeros((), dtype=torch.float32, device=dev)
            for pg in self._groups_of_sharded_params():
                dist.all_reduce(sharded, op=dist.ReduceOp.MAX if inf else dist.ReduceOp.SUM, group=pg)
        parts = [x for x in (sharded, replicated) if x is not None]
        if not parts:
            return None
        dev = parts[0].device
        parts = [x.to(dev) for x in parts]
        total = torch.stack(parts).max() if inf else torch.stack(parts).sum() ** (1.0 / p)
        clip_coef = torch.clamp(self._max_gradient / (total + 1e-6), max=1.0)
        for t in self._sharded_params + self._replicate_params:
            for v in self._local_grads(t):
                v.mul_(clip_coef.to(device=v.device, dtype=v.dtype))
        return total
