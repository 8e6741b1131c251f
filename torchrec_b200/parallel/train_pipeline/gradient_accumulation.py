"""Gradient accumulation on top of any train pipeline (reference train_pipeline/gradient_accumulation.py:31-390).

The wrapped pipeline keeps calling ``optimizer.zero_grad()/step()`` every micro-batch; the wrapper swaps in an optimizer
proxy that only lets every ``num_steps``-th call through, and runs the non-final micro-batches under DDP ``no_sync`` so
the dense all-reduce happens once per accumulation window. Fused (in-backward) embedding optimizers still update per
micro-batch — exactly like the reference's TBE."""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Any, ContextManager, Generic, Iterator, List, Optional, TypeVar

import torch
from torch.nn.parallel import DistributedDataParallel

In = TypeVar("In")
Out = TypeVar("Out")


@dataclass
class GradientAccumulationConfig:
    is_enabled: bool = False
    num_steps: int = 1
    num_warmup_steps: int = 1

    def __post_init__(self) -> None:
        if self.num_steps < 1:
            raise ValueError(f"num_steps must be >= 1, got {self.num_steps}")
        if self.num_warmup_steps < 1:
            raise ValueError(f"num_warmup_steps must be >= 1, got {self.num_warmup_steps}. At least 1 warmup step is required for DDP static_graph compatibility.")
        if self.num_steps > 1 and not self.is_enabled:
            self.is_enabled = True


class _GAOptimizerWrapper:
    """Optimizer proxy: ``zero_grad`` only at the start of a window, ``step`` only at its end."""

    def __init__(self, optimizer: torch.optim.Optimizer, config: GradientAccumulationConfig) -> None:
        self._optimizer = optimizer
        self._config = config
        self._step_count = 0

    def _in_warmup(self) -> bool:
        return self._step_count < self._config.num_warmup_steps

    def _window_pos(self) -> int:
        return (self._step_count - self._config.num_warmup_steps) % self._config.num_steps

    def _should_step(self) -> bool:
        return self._in_warmup() or self._window_pos() == self._config.num_steps - 1

    def zero_grad(self, set_to_none: bool = True) -> None:
        if self._in_warmup() or self._window_pos() == 0:
            self._optimizer.zero_grad(set_to_none=set_to_none) if _accepts_set_to_none(self._optimizer) else self._optimizer.zero_grad()

    def step(self, *args: Any, **kwargs: Any) -> None:
        if self._should_step():
            self._optimizer.step(*args, **kwargs)
        self.advance_step()

    def advance_step(self) -> None:
        self._step_count += 1

    def reset(self) -> None:
        self._step_count = 0

    def set_step(self, step: int) -> None:
        self._step_count = step

    def __getattr__(self, name: str) -> Any:
        return getattr(self._optimizer, name)


def _accepts_set_to_none(opt: Any) -> bool:
    import inspect

    try:
        return "set_to_none" in inspect.signature(opt.zero_grad).parameters
    except (TypeError, ValueError):
        return False


class GradientAccumulationWrapper(Generic[In, Out]):
    def __init__(self, pipeline: Any, optimizer: torch.optim.Optimizer, model: torch.nn.Module, config: GradientAccumulationConfig) -> None:
        self._pipeline = pipeline
        self._model = model
        self._config = config
        self._optimizer_wrapper = _GAOptimizerWrapper(optimizer, config)
        self._cached_ddp_modules: Optional[List[Any]] = None
        if config.is_enabled and hasattr(pipeline, "_optimizer"):
            pipeline._optimizer = self._optimizer_wrapper

    def _should_sync_grad(self, is_last_batch: bool = False) -> bool:
        if not self._config.is_enabled or is_last_batch:
            return True
        return self._optimizer_wrapper._should_step()

    def _get_ddp_modules(self) -> List[Any]:
        if self._cached_ddp_modules is None:
            mods = [m for m in self._model.modules() if isinstance(m, DistributedDataParallel)]
            inner = getattr(self._model, "_dmp_wrapped_module", None)
            if isinstance(inner, DistributedDataParallel) and inner not in mods:
                mods.append(inner)
            self._cached_ddp_modules = mods
        return self._cached_ddp_modules

    def _get_no_sync_context(self) -> ContextManager[None]:
        stack = contextlib.ExitStack()
        for m in self._get_ddp_modules():
            stack.enter_context(m.no_sync())
        return stack

    def progress(self, dataloader_iter: Iterator[In], is_last_batch: bool = False) -> Out:
        if not self._config.is_enabled:
            return self._pipeline.progress(dataloader_iter)
        if self._should_sync_grad(is_last_batch):
            if is_last_batch and not self._optimizer_wrapper._should_step():
                # flush a partial window: make this micro-batch the closing one
                cfg, w = self._config, self._optimizer_wrapper
                w._step_count += cfg.num_steps - 1 - w._window_pos()
            return self._pipeline.progress(dataloader_iter)
        with self._get_no_sync_context():
            return self._pipeline.progress(dataloader_iter)

    def reset(self) -> None:
        self._optimizer_wrapper.reset()
        if hasattr(self._pipeline, "reset"):
            self._pipeline.reset()

    @property
    def optimizer_wrapper(self) -> _GAOptimizerWrapper:
        return self._optimizer_wrapper

    @property
    def current_step(self) -> int:
        return self._optimizer_wrapper._step_count

    def set_step(self, step: int) -> None:
        self._optimizer_wrapper.set_step(step)

    def __getattr__(self, name: str) -> Any:
        return getattr(self._pipeline, name)
