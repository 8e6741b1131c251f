"""Stages of a ``StagedTrainPipeline``: the stage descriptor, the (output, event) pair a stage hands to the next one, and the helper
that turns sharded-module input dists into two stage callables.

Parity: reference ``train_pipeline/pipeline_stage.py`` (``RunnableType`` / ``StageOut`` :60-72, ``PipelineStage`` :74, ``SparseDataDistUtil`` :100-330)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, Generic, Optional, TypeVar

import torch
from torch import nn

from ..types import ShardedModule
from .pipeline_context import TrainPipelineContext

In = TypeVar("In")
RunnableType = Callable[..., Any]
StageOut = Optional[Any]


@dataclass
class StageOutputWithEvent(Generic[In]):
    """A stage's output + the CUDA event recorded when the stage's stream finished producing it."""

    output: Optional[In]
    event: Optional[torch.cuda.Event] = None


@dataclass
class PipelineStage:
    """One user-defined stage of a ``StagedTrainPipeline``: ``runnable(batch) -> batch`` runs on ``stream``; ``fill_callback`` runs once
    per progress() after the stage was launched (e.g. wait for the splits exchange), ``data_exhausted_callback`` when the
    dataloader ran dry while this stage still holds batches."""

    name: str
    runnable: Callable[[Any], Any]
    stream: Optional[torch.Stream] = None
    fill_callback: Optional[Callable[[], None]] = None
    data_exhausted_callback: Optional[Callable[[], None]] = None


class SparseDataDistUtil(Generic[In]):
    """The two sparse-dist stage callables for a ``StagedTrainPipeline``:

        util = SparseDataDistUtil(model, data_dist_stream)
        stages = [PipelineStage("copy", copy_fn, memcpy_stream), PipelineStage("dist", util.start_sparse_data_dist, data_dist_stream,
                                                                             fill_callback=util.wait_sparse_data_dist)]

    ``start_sparse_data_dist(batch)`` launches every sharded module's input dist for ``batch``; ``wait_sparse_data_dist()`` finishes
    them; the model's sharded modules then find their distributed inputs in the context of the batch they are called with."""

    def __init__(self, model: nn.Module, data_dist_stream: Optional[torch.Stream], apply_jit: bool = False, prefetch_stream: Optional[torch.Stream] = None) -> None:
        self.model = model
        self.data_dist_stream = data_dist_stream
        self.prefetch_stream = prefetch_stream
        self.context = TrainPipelineContext(version=1)
        self._modules: Dict[str, ShardedModule] = {n: m for n, m in model.named_modules() if isinstance(m, ShardedModule)}
        self._getters: Dict[str, Callable[[Any], Any]] = {}
        self._original: Dict[str, Callable[..., Any]] = {}
        self._pending: Dict[int, TrainPipelineContext] = {}
        for name, m in self._modules.items():
            self._original[name] = m.forward
            m.forward = self._make_forward(name, m)  # type: ignore[method-assign]

    def _find_kjt(self, batch: Any, module: ShardedModule) -> Any:
        from ...sparse.jagged_tensor import KeyedJaggedTensor

        for attr in ("sparse_features", "id_list_features", "features"):
            v = getattr(batch, attr, None)
            if isinstance(v, KeyedJaggedTensor):
                return v
        if isinstance(batch, KeyedJaggedTensor):
            return batch
        raise RuntimeError(f"SparseDataDistUtil: cannot find the KeyedJaggedTensor input in a {type(batch).__name__}")

    def _make_forward(self, name: str, module: ShardedModule) -> Callable[..., Any]:
        def fwd(*input: Any, **kwargs: Any) -> Any:
            ctx = self._pending.pop(id(input[0]), None) if input else None
            if ctx is None or name not in ctx.input_dist_tensors_requests:
                if ctx is not None:
                    self._pending[id(input[0])] = ctx
                return self._original[name](*input, **kwargs)
            request = ctx.input_dist_tensors_requests.pop(name)
            mctx = ctx.module_contexts.pop(name)
            if ctx.input_dist_tensors_requests:
                self._pending[id(input[0])] = ctx
            if self.data_dist_stream is not None:
                with torch.cuda.stream(self.data_dist_stream):
                    data = request.wait()
                torch.cuda.current_stream().wait_stream(self.data_dist_stream)
            else:
                data = request.wait()
            return module.compute_and_output_dist(mctx, data)

        return fwd

    def start_sparse_data_dist(self, batch: In) -> In:
        ctx = TrainPipelineContext(version=1)
        for name, m in self._modules.items():
            kjt = self._find_kjt(batch, m)
            mctx = m.create_context()
            ctx.module_contexts[name] = mctx
            ctx.input_dist_splits_requests[name] = m.input_dist(mctx, kjt)
            self._pending[id(kjt)] = ctx
        self.context = ctx
        return batch

    def wait_sparse_data_dist(self) -> None:
        ctx = self.context
        for name, aw in list(ctx.input_dist_splits_requests.items()):
            ctx.input_dist_tensors_requests[name] = aw.wait()
        ctx.input_dist_splits_requests.clear()

    def detach(self) -> nn.Module:
        for name, m in self._modules.items():
            m.forward = self._original[name]  # type: ignore[method-assign]
        self._pending.clear()
        return self.model
