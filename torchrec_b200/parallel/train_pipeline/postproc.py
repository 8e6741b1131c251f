"""Reference import path ``torchrec/distributed/train_pipeline/postproc.py`` (``NoOpStream`` :23, ``PipelinedPostproc`` :37); implementation in ``utils.py``."""
from __future__ import annotations

import queue
import threading
from typing import Any, Callable, Dict, Generic, Iterator, List, Optional, TypeVar
import torch
from torch import nn
from torch.autograd.profiler import record_function
from ..types import ShardedModule
from .pipeline_context import TrainPipelineContext


class NoOpStream:
    """Stand-in for a CUDA stream on CPU: a context manager whose stream operations do nothing."""

    def __init__(self, stream: Any = None) -> None:
        self._stream = stream

    def __enter__(self) -> "NoOpStream":
        return self

    def __exit__(self, *exc: Any) -> None:
        return None

    def wait_stream(self, other: Any) -> None:
        pass

    def synchronize(self) -> None:
        pass

    def record_event(self, event: Any = None) -> Any:
        return event


class PipelinedPostproc(nn.Module):
    """A batch post-processing module (id remapping, feature crossing, ...) hoisted out of the model's forward into the pipeline's
    data-dist stage: the pipeline calls it once per batch on the data-dist stream, the model's own call returns the cached result
    for that batch context (so it runs a step early and only once even if several sharded modules consume its output)."""

    def __init__(self, postproc_module: nn.Module, fqn: str, context: Optional[TrainPipelineContext] = None, default_stream: Optional[torch.Stream] = None,
                 dist_stream: Optional[torch.Stream] = None) -> None:
        super().__init__()
        self._postproc_module = postproc_module
        self._fqn = fqn
        self._context = context
        self._default_stream = default_stream
        self._dist_stream = dist_stream

    @property
    def postproc_module(self) -> nn.Module:
        return self._postproc_module

    @property
    def fqn(self) -> str:
        return self._fqn

    def set_context(self, context: TrainPipelineContext) -> None:
        self._context = context

    def forward(self, *input: Any, **kwargs: Any) -> Any:
        ctx = self._context
        cache: Optional[Dict[str, Any]] = getattr(ctx, "postproc_fwd_results", None) if ctx is not None else None
        if cache is not None and self._fqn in cache:
            res = cache[self._fqn]
            if self._dist_stream is not None and self._default_stream is not None:
                self._default_stream.wait_stream(self._dist_stream)  # result was produced on the data-dist stream
            return res
        with record_function(f"## pipelined_postproc {self._fqn} ##"):
            res = self._postproc_module(*input, **kwargs)
        if ctx is not None:
            if cache is None:
                cache = {}
                ctx.postproc_fwd_results = cache  # type: ignore[attr-defined]
            cache[self._fqn] = res
        return res
