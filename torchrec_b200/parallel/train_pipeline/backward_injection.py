"""Backward injection: run extra work (prefetch, stash, communication) at a chosen point of the backward pass
(reference train_pipeline/backward_injection.py:103-460)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum, unique
from typing import Any, Callable, Iterator, List, Optional, Protocol, runtime_checkable

import torch
from torch import nn


BackwardHookWork = Callable[[Any], None]  # work injected into the backward pass; receives the train pipeline


@unique
class InjectionTargetType(Enum):
    PARAM_GRAD = "param_grad"   # post-accumulate-grad hook on one parameter of the target module
    ACTIVATION = "activation"   # forward hook finds an output tensor, tensor hook fires when its gradient arrives


@runtime_checkable
class GradTensorFinder(Protocol):
    def __call__(self, module: nn.Module, inputs: Any, output: Any) -> Optional[torch.Tensor]:
        ...


class FirstGradTensorFinder:
    """First tensor requiring grad found in the module output (awaitables, KeyedTensors, dicts, tuples are searched)."""

    def _search(self, data: Any) -> Optional[torch.Tensor]:
        if isinstance(data, torch.Tensor):
            return data if data.requires_grad else None
        if hasattr(data, "wait") and hasattr(data, "_wait_impl"):
            return self._search(data.wait())
        if hasattr(data, "values") and callable(data.values) and not isinstance(data, dict):
            try:
                return self._search(data.values())
            except TypeError:
                return None
        if isinstance(data, dict):
            data = list(data.values())
        if isinstance(data, (list, tuple)):
            for x in data:
                t = self._search(x)
                if t is not None:
                    return t
        return None

    def __call__(self, module: nn.Module, inputs: Any, output: Any) -> Optional[torch.Tensor]:
        return self._search(output)


@dataclass
class OutputDistTensorFinder:
    """Hook point at the output dist of ONE sharding type of a pipelined sharded embedding module: the tensor whose gradient arrives when
    the backward of that sharding type's output collective starts. The module output is an awaitable made of per-sharding awaitables
    (``_awaitables`` / ``_awaitables_per_sharding`` + ``_sharding_types``); the one of ``sharding_type`` is picked (replicated tables
    have no collective and are skipped) and the first gradient-carrying tensor of its result returned - None when that sharding type
    produces no differentiable output on this rank."""

    sharding_type: Any = None

    def __post_init__(self) -> None:
        from ..types import ShardingType

        if self.sharding_type is None:
            self.sharding_type = ShardingType.TABLE_WISE

    def __call__(self, module: nn.Module, inputs: Any, output: Any) -> Optional[torch.Tensor]:
        from ..types import NoWait, ShardingType

        if isinstance(output, tuple):  # managed-collision modules return (embeddings, remapped features)
            output = output[0]
        awaitables = getattr(output, "_awaitables", None)
        if awaitables is None:
            awaitables = getattr(output, "_awaitables_per_sharding", None)
        types = getattr(output, "_sharding_types", None)
        if awaitables is None or types is None:
            # this framework's fused engine issues ONE output collective for all sharding types: the first gradient tensor of the
            # module output is the hook point of every type
            return FirstGradTensorFinder()(module, inputs, output)
        for w, st in zip(awaitables, types):
            if isinstance(w, NoWait):
                continue
            if ShardingType(st) == self.sharding_type:
                dummy = getattr(getattr(w, "_tensor_awaitable", None), "dummy_tensor", None)
                return dummy if dummy is not None else FirstGradTensorFinder()._search(w)
        raise RuntimeError(f"Could not find awaitable for sharding type: {self.sharding_type}")


@dataclass
class InjectionSite:
    fqn: str
    tensor_finder: GradTensorFinder = FirstGradTensorFinder()
    target_type: InjectionTargetType = InjectionTargetType.ACTIVATION
    hook_position: float = 1.0


class _Handles:
    def __init__(self, handles: List[Any]) -> None:
        self._handles = handles

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def will_hook_fire(p: torch.Tensor) -> bool:
    """A post-accumulate-grad hook only fires for leaf parameters that autograd accumulates into."""
    return isinstance(p, torch.Tensor) and p.requires_grad and p.is_leaf and not getattr(p, "_in_backward_optimizers", None)


def _walk_outward(start: int, n: int) -> Iterator[int]:
    yield start
    for d in range(1, n):
        if start - d >= 0:
            yield start - d
        if start + d < n:
            yield start + d


def _position_to_index(position: float, length: int) -> int:
    return min(max(int(round(position * (length - 1))), 0), length - 1)


def _register_param_grad_hook(site: InjectionSite, target: nn.Module, hook_fn: Callable[[torch.Tensor], None]):
    params = list(target.parameters())
    if not params:
        raise ValueError(f"register_backward_hook: module '{site.fqn}' has no parameters.")
    start = _position_to_index(site.hook_position, len(params))
    for i in _walk_outward(start, len(params)):
        if will_hook_fire(params[i]):
            p = params[i]

            def _grad_adapter(param: torch.Tensor) -> None:
                hook_fn(param.grad if param.grad is not None else param)

            return p.register_post_accumulate_grad_hook(_grad_adapter)
    raise ValueError(f"register_backward_hook: module '{site.fqn}' has no parameter whose gradient hook would fire.")


def _register_activation_hook(site: InjectionSite, target: nn.Module, hook_fn: Callable[[torch.Tensor], None]):
    inner: List[Any] = []

    def fwd_hook(module: nn.Module, inputs: Any, output: Any) -> None:
        if not torch.is_grad_enabled():
            return
        t = site.tensor_finder(module, inputs, output)
        if t is None:
            raise RuntimeError(f"backward injection at '{site.fqn}': tensor_finder found no tensor requiring grad in the output")
        inner.append(t.register_hook(lambda g: (hook_fn(g), None)[1]))

    h = target.register_forward_hook(fwd_hook)
    return _Handles([h])


def register_backward_hook(site: InjectionSite, model: nn.Module, hook_fn: Callable[[torch.Tensor], None]):
    try:
        target = model.get_submodule(site.fqn)
    except AttributeError:
        raise ValueError(f"register_backward_hook: module '{site.fqn}' not found in model.")
    if site.target_type == InjectionTargetType.PARAM_GRAD:
        return _register_param_grad_hook(site, target, hook_fn)
    if site.target_type == InjectionTargetType.ACTIVATION:
        return _register_activation_hook(site, target, hook_fn)
    raise ValueError(f"register_backward_hook: unknown target_type '{site.target_type}'.")
