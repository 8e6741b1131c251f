"""Pipeline building blocks that live outside the pipeline classes: background data loading, pipelined batch post-processing and the
sparse-data-dist helper for ``StagedTrainPipeline``.

Parity: reference ``train_pipeline/utils.py`` (``DataLoadingThread`` :780-900), ``postproc.py`` (``PipelinedPostproc`` :37-300),
``pipeline_stage.py`` (``SparseDataDistUtil`` :100-330)."""
from __future__ import annotations

import queue
import contextlib
import threading
from collections import deque
from concurrent.futures import Future
from typing import Any, Callable, Dict, Generator, Generic, Iterator, List, Optional, TypeVar

import torch
from torch import nn
from torch.autograd.profiler import record_function

from ..types import ShardedModule
from .pipeline_context import TrainPipelineContext

In = TypeVar("In")


def _to_device(batch: Any, device: torch.device, non_blocking: bool = True) -> Any:
    return batch.to(device=device, non_blocking=non_blocking) if hasattr(batch, "to") else batch


class DataLoadingThread(threading.Thread, Generic[In]):
    """Pulls batches from the dataloader iterator on a background thread, starts their H2D copy on ``memcpy_stream`` and hands
    ``(batch, copy_done_event)`` to the trainer through a bounded queue. The host work of ``next(dataloader)`` (collation,
    pin-memory) then overlaps the trainer's kernel launches instead of sitting in front of them."""

    def __init__(self, device: torch.device, dataloader_iter: Iterator[In], to_device_non_blocking: bool = True, memcpy_stream_priority: int = 0,
                 memcpy_stream: Optional[torch.Stream] = None, queue_size: int = 2) -> None:
        super().__init__(daemon=True)
        self._device = torch.device(device)
        self._iter = dataloader_iter
        self._non_blocking = to_device_non_blocking
        self._queue: "queue.Queue[Any]" = queue.Queue(maxsize=max(1, queue_size))
        self._stop_event = threading.Event()
        self._exc: Optional[BaseException] = None
        if memcpy_stream is None and self._device.type == "cuda":
            memcpy_stream = torch.cuda.Stream(device=self._device, priority=memcpy_stream_priority)
        self._memcpy_stream = memcpy_stream

    def run(self) -> None:
        if self._device.type == "cuda":
            torch.cuda.set_device(self._device)
        try:
            while not self._stop_event.is_set():
                with record_function("## load_batch ##"):
                    batch = next(self._iter, None)
                if batch is None:
                    break
                ev = None
                if self._memcpy_stream is not None:
                    with torch.cuda.stream(self._memcpy_stream):
                        batch = _to_device(batch, self._device, self._non_blocking)
                        ev = torch.cuda.Event()
                        ev.record(self._memcpy_stream)
                else:
                    batch = _to_device(batch, self._device, self._non_blocking)
                while not self._stop_event.is_set():
                    try:
                        self._queue.put((batch, ev), timeout=0.05)
                        break
                    except queue.Full:
                        continue
        except BaseException as e:  # surfaced to the trainer by get()
            self._exc = e
        finally:
            while True:
                try:
                    self._queue.put(None, timeout=0.05)
                    break
                except queue.Full:
                    if self._stop_event.is_set():
                        break

    def stop(self) -> None:
        self._stop_event.set()

    def get(self, timeout: Optional[float] = None) -> Optional[In]:
        """Next device batch (the current stream is made to wait for its copy), or None at the end of the data."""
        item = self._queue.get(timeout=timeout)
        if item is None:
            self._queue.put(None)  # keep signalling the end to later calls
            if self._exc is not None:
                raise self._exc
            return None
        batch, ev = item
        if ev is not None:
            torch.cuda.current_stream(self._device).wait_event(ev)
            if hasattr(batch, "record_stream"):
                batch.record_stream(torch.cuda.current_stream(self._device))
        return batch






def get_h2d_func(batch: Any, device: torch.device) -> Any:
    """The default host-to-device move of a pipeline stage: ``batch.to(device, non_blocking=True)``."""
    return batch.to(device, non_blocking=True)


class FutureDeque(deque):
    """A deque whose entries may be ``concurrent.futures.Future``s (batches still being produced by a worker): reading an entry
    (index, ``pop``, ``popleft``) waits for it and, for an index read, replaces the future by its result."""

    def __getitem__(self, index: Any) -> Any:
        item = super().__getitem__(index)
        if isinstance(item, Future):
            item = item.result()
            super().__setitem__(index, item)
        return item

    def pop(self) -> Any:  # type: ignore[override]
        item = super().pop()
        return item.result() if isinstance(item, Future) else item

    def popleft(self) -> Any:
        item = super().popleft()
        return item.result() if isinstance(item, Future) else item


@contextlib.contextmanager
def use_context_for_postprocs(pipelined_postprocs: List[Any], next_batch_context: Any) -> Generator[None, None, None]:
    """While the block runs the pipelined post-processing modules write into the context of the NEXT batch (their outputs are cached
    there for its forward); the contexts of the current batch are restored afterwards."""
    original = [p.get_context() for p in pipelined_postprocs]
    for p in pipelined_postprocs:
        p.set_context(next_batch_context)
    try:
        yield
    finally:
        for p, ctx in zip(pipelined_postprocs, original):
            p.set_context(ctx)


def prefetch_embeddings(context: Any, pipelined_modules: List[Any], device: torch.device, stream_context: Callable[[Optional[torch.Stream]], Any],
                        data_dist_stream: Optional[torch.Stream], forward_stream: Optional[torch.Stream]) -> None:
    """For every pipelined sharded module: finish the input dist of the batch (on the data-dist stream), let the prefetch stream wait
    for it, start the cache prefetch of the ids and park input + module context in ``context.module_input_post_prefetch`` /
    ``module_contexts_post_prefetch`` for the forward (reference train_pipeline/utils.py:728)."""
    if data_dist_stream is None:
        return
    cur = torch.get_device_module(device).current_stream() if device.type == "cuda" else None
    for m in pipelined_modules:
        name = getattr(m.forward, "_name", None) or getattr(m.forward, "name", None)
        assert name in context.input_dist_tensors_requests, f"no input dist request for {name}"
        request = context.input_dist_tensors_requests.pop(name)
        with stream_context(data_dist_stream):
            dist_input = request.wait()
        if cur is not None:
            cur.wait_stream(data_dist_stream)
            if hasattr(dist_input, "record_stream"):
                dist_input.record_stream(cur)
        mctx = context.module_contexts_next_batch.pop(name, None) or context.module_contexts.get(name)
        if hasattr(m, "prefetch"):
            m.prefetch(ctx=mctx, dist_input=dist_input, forward_stream=forward_stream)
        context.module_input_post_prefetch[name] = dist_input
        context.module_contexts_post_prefetch[name] = mctx


# ---- moved to ``postproc.py`` (their reference import path); still importable from here ----
_MOVED_TO_POSTPROC = ('PipelinedPostproc',)


def __getattr__(name: str):
    if name in _MOVED_TO_POSTPROC:
        from . import postproc as _m

        return getattr(_m, name)
    if name == 'SparseDataDistUtil':
        from . import pipeline_stage as _m

        return _m.SparseDataDistUtil
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
