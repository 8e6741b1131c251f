"""Replacement ``forward`` callables that the pipelines install on sharded modules (reference train_pipeline/runtime_forwards.py:54-456).

All of them answer the same question at the model's call site - "where is the already distributed input / the already computed
embedding of THIS batch?" - from the batch's ``TrainPipelineContext``:

* ``PipelinedForward``            input dist done ahead  -> run ``compute_and_output_dist`` now
* ``EmbeddingPipelinedForward``   lookup + output dist started ahead (semi-sync / fused pipelines) -> wait for it
* ``InSyncEmbeddingPipelinedForward``  same, but never returns embeddings computed before the last optimizer step
* ``CPUEmbeddingPipelinedForward``     embeddings computed by a host-side model copy, moved to the device here
* ``PrefetchPipelinedForward``    input went through the cache-prefetch stage
* ``KJTAllToAllForward``          splits-fusing replacement of ``KJTAllToAll.forward``"""
from __future__ import annotations

from typing import Any, Callable, Generic, List, Optional, TypeVar

import torch
from torch.autograd.profiler import record_function

from ..types import Awaitable, ShardedModule
from .pipeline_context import EmbeddingTrainPipelineContext, PrefetchTrainPipelineContext, TrainPipelineContext
from .train_pipelines import PipelinedForward  # noqa: F401  (same class the pipelines use)
from .types import CallArgs

TForwardContext = TypeVar("TForwardContext", bound=TrainPipelineContext)


class BaseForward(Generic[TForwardContext]):
    def __init__(self, name: str, args: Optional[CallArgs], module: ShardedModule, context: TForwardContext, stream: Optional[torch.Stream] = None) -> None:
        self._name = name
        self._args = args
        self._module = module
        self._context = context
        self._stream = stream
        self._device: Optional[torch.device] = stream.device if stream is not None else None

    @property
    def name(self) -> str:
        return self._name

    @property
    def args(self) -> Optional[CallArgs]:
        return self._args

    def set_context(self, context: TForwardContext) -> None:
        self._context = context

    def get_context(self) -> TForwardContext:
        return self._context

    def _sync_stream(self, *objs: Any) -> None:
        if self._stream is None:
            return
        cur = torch.get_device_module(self._stream.device).current_stream()
        cur.wait_stream(self._stream)
        for o in objs:
            if hasattr(o, "record_stream"):
                o.record_stream(cur)


class EmbeddingPipelinedForward(BaseForward[EmbeddingTrainPipelineContext]):
    """The pipeline already called ``compute_and_output_dist`` for this batch and stored the awaitable in
    ``context.embedding_a2a_requests[name]``: wait for it (on the embedding stream) and hand the result to the model."""

    def __call__(self, *input: Any, **kwargs: Any) -> Any:
        assert self._name in self._context.embedding_a2a_requests, f"no precomputed embedding for {self._name}: was the lookup stage run for this batch?"
        aw = self._context.embedding_a2a_requests.pop(self._name)
        with record_function("## wait_embedding_a2a ##"):
            if self._stream is not None:
                with torch.get_device_module(self._stream.device).stream(self._stream):
                    res = aw.wait() if isinstance(aw, Awaitable) or hasattr(aw, "wait") else aw
            else:
                res = aw.wait() if hasattr(aw, "wait") else aw
        self._sync_stream(res)
        self._context.module_contexts.pop(self._name, None)
        return res


class InSyncEmbeddingPipelinedForward(EmbeddingPipelinedForward):
    """Like ``EmbeddingPipelinedForward`` but refuses stale results: if the precomputed embedding was produced before the weights'
    latest update (``context.index`` older than ``fresh_since``) the lookup is redone now from the distributed input."""

    def __init__(self, *args: Any, fresh_since: Callable[[], int] = lambda: -1, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._fresh_since = fresh_since

    def __call__(self, *input: Any, **kwargs: Any) -> Any:
        idx = self._context.index if self._context.index is not None else 0
        if idx < self._fresh_since() and self._name in self._context.input_dist_tensors_requests:
            self._context.embedding_a2a_requests.pop(self._name, None)
            data = self._context.input_dist_tensors_requests.pop(self._name).wait()
            return self._module.compute_and_output_dist(self._context.module_contexts.pop(self._name), data)
        return super().__call__(*input, **kwargs)


class CPUEmbeddingPipelinedForward(EmbeddingPipelinedForward):
    """Embeddings were computed by a CPU copy of the sparse modules (host-offloaded eval / hybrid pipelines): move them to the
    training device at the call site."""

    def __init__(self, *args: Any, device: Optional[torch.device] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._target = torch.device(device) if device is not None else None

    def __call__(self, *input: Any, **kwargs: Any) -> Any:
        res = super().__call__(*input, **kwargs)
        if self._target is None:
            return res
        if hasattr(res, "to"):
            return res.to(self._target, non_blocking=True)
        if isinstance(res, dict):
            return {k: v.to(self._target, non_blocking=True) for k, v in res.items()}
        return res


class PrefetchPipelinedForward(BaseForward[PrefetchTrainPipelineContext]):
    """The distributed input went through the prefetch stage (cache rows resident): take it from the post-prefetch slots."""

    def __init__(self, name: str, args: Optional[CallArgs], module: ShardedModule, context: PrefetchTrainPipelineContext, prefetch_stream: Optional[torch.Stream] = None) -> None:
        super().__init__(name, args, module, context, prefetch_stream)

    def __call__(self, *input: Any, **kwargs: Any) -> Any:
        assert self._name in self._context.module_input_post_prefetch, f"{self._name}: prefetch stage did not run for this batch"
        data = self._context.module_input_post_prefetch.pop(self._name)
        mctx = self._context.module_contexts_post_prefetch.pop(self._name)
        self._sync_stream(data, mctx)
        return self._module.compute_and_output_dist(mctx, data)


class PrefetchEmbeddingPipelinedForward(PrefetchPipelinedForward):
    """Prefetch + early lookup: the embedding awaitable is in ``embedding_a2a_requests`` (contexts that carry both)."""

    def __call__(self, *input: Any, **kwargs: Any) -> Any:
        reqs = getattr(self._context, "embedding_a2a_requests", None)
        if reqs and self._name in reqs:
            aw = reqs.pop(self._name)
            res = aw.wait() if hasattr(aw, "wait") else aw
            self._sync_stream(res)
            self._context.module_input_post_prefetch.pop(self._name, None)
            self._context.module_contexts_post_prefetch.pop(self._name, None)
            return res
        return super().__call__(*input, **kwargs)


class KJTAllToAllForward:
    """``KJTAllToAll.forward`` replacement used when several modules' split exchanges are fused into one collective: it returns the
    splits awaitable without launching its own size exchange (``fused_splits`` fills the sizes in later)."""

    def __init__(self, pg: Any, splits: List[int], stagger: int = 1) -> None:
        self._pg, self._splits, self._stagger = pg, splits, stagger

    def __call__(self, kjt: Any) -> Any:
        from ..dist_data import KJTAllToAllSplitsAwaitable

        with record_function("## all2all_data:kjt splits (fused) ##"):
            return KJTAllToAllSplitsAwaitable(pg=self._pg, input=kjt, splits=self._splits, labels=kjt.dist_labels(), tensor_splits=kjt.dist_splits(self._splits),
                                              input_tensors=kjt.dist_tensors(), keys=kjt.keys(), device=kjt.device(), stagger=self._stagger)
