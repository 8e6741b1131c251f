"""Fused-optimizer protocol (reference torchrec/optim/fused.py): the optimizer step already
happened inside the embedding backward kernel; ``step()`` only pushes hyper-parameters."""
from __future__ import annotations

import abc
from typing import Any

from torch import optim

from .keyed import KeyedOptimizer


class FusedOptimizer(KeyedOptimizer, abc.ABC):
    @abc.abstractmethod
    def step(self, closure: Any = None) -> None:
        ...

    @abc.abstractmethod
    def zero_grad(self, set_to_none: bool = False) -> None:
        ...

    def __repr__(self) -> str:
        return optim.Optimizer.__repr__(self)


class EmptyFusedOptimizer(FusedOptimizer):
    def __init__(self) -> None:
        super().__init__({}, {}, {})

    def step(self, closure: Any = None) -> None:
        pass

    def zero_grad(self, set_to_none: bool = False) -> None:
        pass


class FusedOptimizerModule(abc.ABC):
    """Module that owns a fused optimizer."""

    @property
    @abc.abstractmethod
    def fused_optimizer(self) -> KeyedOptimizer:
        ...
