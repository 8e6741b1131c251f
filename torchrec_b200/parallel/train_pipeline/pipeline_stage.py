"""Reference import path ``torchrec/distributed/train_pipeline/pipeline_stage.py`` (``RunnableType`` / ``StageOut`` :60-72, ``PipelineStage`` :74,
``SparseDataDistUtil`` :100); implementations in ``train_pipelines.py`` / ``utils.py``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Generic, Optional, TypeVar

import torch

from .train_pipelines import PipelineStage  # noqa: F401
from .utils import SparseDataDistUtil  # noqa: F401

In = TypeVar("In")
RunnableType = Callable[..., Any]


@dataclass
class StageOutputWithEvent(Generic[In]):
    """A stage's output + the CUDA event recorded when the stage's stream finished producing it."""

    output: Optional[In]
    event: Optional[torch.cuda.Event] = None


StageOut = Optional[Any]
