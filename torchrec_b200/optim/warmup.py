"""Stage-wise learning-rate schedules applied on top of any keyed optimizer.

A schedule is a list of ``WarmupStage``s; stage k is active until the global step counter passes its ``max_iters`` (an
absolute step), after the last stage the multiplier stays at 1. The multiplier of a step is looked up in a table of
shape functions (``SCHEDULE_SHAPES``), so adding a policy is one entry. The step counter and the active stage travel in
the optimizer state under a pseudo parameter (``__warmup``): resuming from a checkpoint restores the exact learning rate.

Capability parity: reference ``torchrec/optim/warmup.py`` (policies incl. INTERPOLATE ``:23-31``, stage fields ``:34-48``,
state handling ``:114-190``).
"""
from __future__ import annotations

import logging
import math
from dataclasses import dataclass
from enum import Enum, unique
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch

from .keyed import KeyedOptimizer, OptimizerWrapper

logger = logging.getLogger(__name__)

_FOREVER = 1 << 63


@unique
class WarmupPolicy(Enum):
    NONE = "none"
    LINEAR = "linear"
    CONSTANT = "constant"
    POLY = "poly"
    STEP = "step"
    INVSQRT = "inv_sqrt"
    COSINE_ANNEALING_WARM_RESTARTS = "cosine_annealing_warm_restarts"
    INTERPOLATE = "interpolate"


@dataclass
class WarmupStage:
    policy: WarmupPolicy = WarmupPolicy.LINEAR
    max_iters: int = 1            # absolute step at which the stage ends
    value: float = 1.0            # start multiplier (LINEAR / INTERPOLATE), constant, exponent (POLY), decay factor (STEP), floor (COSINE)
    lr_scale: float = 1.0         # extra factor on top of the shape
    decay_iters: int = -1         # POLY horizon / STEP interval (defaults: max_iters / 1)
    sgdr_period: int = 1          # COSINE restart period
    start_interpolating_iters: Optional[int] = None  # INTERPOLATE: step at which the multiplier is `value` ...
    end_value: Optional[float] = None                # ... and the multiplier reached at `max_iters`


def _shape_linear(s: WarmupStage, it: int) -> float:
    return s.value + (1.0 - s.value) * (it / s.max_iters)


def _shape_poly(s: WarmupStage, it: int) -> float:
    return (1.0 - it / s.decay_iters) ** s.value


def _shape_step(s: WarmupStage, it: int) -> float:
    return s.value ** (it // s.decay_iters)


def _shape_cosine(s: WarmupStage, it: int) -> float:
    phase = (it % s.sgdr_period) / s.sgdr_period
    return s.value + (1.0 - s.value) * 0.5 * (1.0 + math.cos(math.pi * phase))


def _shape_interpolate(s: WarmupStage, it: int) -> float:
    t0 = float(s.start_interpolating_iters)  # type: ignore[arg-type]
    frac = (it - t0) / (s.max_iters - t0)
    return s.value + (float(s.end_value) - s.value) * frac  # type: ignore[arg-type]


SCHEDULE_SHAPES: Dict[WarmupPolicy, Callable[[WarmupStage, int], float]] = {
    WarmupPolicy.NONE: lambda s, it: 1.0,
    WarmupPolicy.LINEAR: _shape_linear,
    WarmupPolicy.CONSTANT: lambda s, it: s.value,
    WarmupPolicy.POLY: _shape_poly,
    WarmupPolicy.STEP: _shape_step,
    WarmupPolicy.INVSQRT: lambda s, it: 1.0 / math.sqrt(max(it, 1)),
    WarmupPolicy.COSINE_ANNEALING_WARM_RESTARTS: _shape_cosine,
    WarmupPolicy.INTERPOLATE: _shape_interpolate,
}


def lr_multiplier(stage: WarmupStage, it: int) -> float:
    return SCHEDULE_SHAPES[stage.policy](stage, it) * stage.lr_scale


def _validated(stages: List[WarmupStage]) -> List[WarmupStage]:
    """Check that stage ends increase, fill per-policy defaults, append the terminal (multiplier 1) stage."""
    prev_end = 0
    for i, st in enumerate(stages):
        if st.max_iters <= prev_end:
            raise AssertionError(f"warm-up stage {i} ends at step {st.max_iters}, not after the previous stage's end ({prev_end})")
        if st.policy is WarmupPolicy.INTERPOLATE:
            if st.start_interpolating_iters is None or st.end_value is None:
                raise AssertionError("INTERPOLATE stages need start_interpolating_iters and end_value")
            if st.max_iters <= st.start_interpolating_iters:
                raise AssertionError(f"INTERPOLATE stage {i}: max_iters {st.max_iters} must exceed start_interpolating_iters {st.start_interpolating_iters}")
        if st.decay_iters <= 0:
            st.decay_iters = 1 if st.policy is WarmupPolicy.STEP else st.max_iters
        prev_end = st.max_iters
    return list(stages) + [WarmupStage(policy=WarmupPolicy.NONE, max_iters=_FOREVER, value=1.0)]


# names kept for callers written against the reference module
_lr_stages = _validated
_get_multiplier = lr_multiplier


class WarmupOptimizer(OptimizerWrapper):
    """Sets ``param_group[lr_param] = lr * multiplier(step)`` on the wrapped optimizer after every step."""

    STATE_KEY = "warmup"

    def __init__(self, optimizer: KeyedOptimizer, stages: List[WarmupStage], lr: float = 0.1, lr_param: str = "lr", param_name: str = "__warmup") -> None:
        super().__init__(optimizer)
        self._stages = _validated(stages)
        self._lr_param = lr_param
        self._lr = lr
        # a pseudo parameter owns the (step, stage) counters so that they are saved / restored with the optimizer state
        self._warmup_param = torch.nn.Parameter()
        self.params[param_name] = self._warmup_param  # type: ignore[index]
        self._apply(0, 0)

    # ---- schedule position ------------------------------------------------------------------------------
    def _position(self) -> Tuple[int, int]:
        entry = self.state.get(self._warmup_param) if hasattr(self.state, "get") else None
        if entry is None:
            return 0, 0
        it, stage = entry[self.STATE_KEY].tolist()
        return int(it), int(stage)

    def _apply(self, it: int, stage: int) -> None:
        lr = self._lr * lr_multiplier(self._stages[stage], it)
        for group in self.param_groups:
            group[self._lr_param] = lr

    # kept for code that drives the schedule directly
    def _set_lr(self, iter_: int, stage_id: int) -> None:
        self._apply(iter_, stage_id)

    def _get_warmup_state(self) -> Tuple[int, int]:
        return self._position()

    # ---- optimizer protocol ----------------------------------------------------------------------------------
    def post_load_state_dict(self) -> None:
        it, stage = self._position()
        logger.info(f"warm-up schedule resumed at step {it} (stage {stage})")
        self._apply(it, stage)

    def step(self, closure: Any = None) -> None:
        super().step(closure)
        it, stage = self._position()
        it += 1
        while it > self._stages[stage].max_iters and stage + 1 < len(self._stages):
            stage += 1
            logger.info(f"warm-up schedule: step {it} enters stage {stage} ({self._stages[stage].policy.value})")
        self._apply(it, stage)
        self.state[self._warmup_param] = {self.STATE_KEY: torch.tensor([it, stage], dtype=torch.long)}  # type: ignore[index]
