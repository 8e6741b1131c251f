"""Small value types of the pipeline package (reference train_pipeline/types.py:15-120)."""
from __future__ import annotations

import abc
from dataclasses import dataclass, field
from enum import Enum, unique
from typing import Any, Dict, List, Optional


class BaseArgInfoStep(abc.ABC):
    """One step of the recipe that extracts a sharded module's input from the batch."""

    @abc.abstractmethod
    def process(self, arg: Any) -> Any:
        ...


@dataclass
class GetAttrArgInfoStep(BaseArgInfoStep):
    attr_name: str

    def process(self, arg: Any) -> Any:
        return getattr(arg, self.attr_name)


@dataclass
class GetItemArgInfoStep(BaseArgInfoStep):
    item_index: Any

    def process(self, arg: Any) -> Any:
        return arg[self.item_index]


@dataclass
class PostprocArgInfoStep(BaseArgInfoStep):
    postproc_module: Any

    def process(self, arg: Any) -> Any:
        return self.postproc_module(arg)


@dataclass
class ArgInfo:
    """How to obtain ONE argument of a pipelined module from the batch: apply ``steps`` in order (attribute access, indexing,
    pipelined post-processing). The hook-based discovery of this framework produces the same recipes as the reference's fx trace
    (``KJTGetter`` paths), this is their declarative form."""

    steps: List[BaseArgInfoStep] = field(default_factory=list)

    def process(self, batch: Any) -> Any:
        arg = batch
        for s in self.steps:
            arg = s.process(arg)
        return arg

    @staticmethod
    def from_path(path: List[Any]) -> "ArgInfo":
        """``[("attr", "sparse_features"), ("item", 0)]`` -> ArgInfo (the format ``KJTGetter`` stores)."""
        steps: List[BaseArgInfoStep] = []
        for kind, key in path:
            steps.append(GetAttrArgInfoStep(key) if kind == "attr" else GetItemArgInfoStep(key))
        return ArgInfo(steps)


@dataclass
class CallArgs:
    args: List[ArgInfo] = field(default_factory=list)
    kwargs: Dict[str, ArgInfo] = field(default_factory=dict)

    def build_args_kwargs(self, batch: Any):
        return [a.process(batch) for a in self.args], {k: v.process(batch) for k, v in self.kwargs.items()}


@unique
class PipelineState(Enum):
    IDLE = 0
    CALL_FWD = 1
    CALL_BWD = 2

    def __str__(self) -> str:
        return self.name


@unique
class PipelinePhase(Enum):
    """Named points of one pipeline step (used by backward-injection work items and trace annotations)."""

    PROGRESS_START = "progress_start"
    COPY_BATCH_TO_GPU = "copy_batch_to_gpu"
    START_SPARSE_DATA_DIST = "start_sparse_data_dist"
    WAIT_SPARSE_DATA_DIST = "wait_sparse_data_dist"
    FORWARD = "forward"
    BACKWARD = "backward"
    OPTIMIZER = "optimizer"
    PROGRESS_END = "progress_end"
