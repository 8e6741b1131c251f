"""Experimental pipeline variants (reference train_pipeline/experimental_pipelines.py:64-1250):

* ``TrainEvalHybridPipelineBase`` — one pipeline object that serves interleaved train and eval batches (``batch.is_eval``
  or an explicit predicate selects the mode) without draining the prefetch queues between them.
* ``EvalPipelineCPUSparse`` — eval with CPU-resident embeddings (HybridEvalDMP): sparse lookup on the host into pinned
  buffers, one H2D copy of pooled embeddings + dense features, dense forward on the GPU.
* ``TrainPipelineSparseDistT`` — the H2D batch copy is issued from a helper thread so Python never blocks on it.
* ``TrainPipelineSparseDistBwdOpt`` — work injected into the backward pass at a chosen module (backward_injection).
* ``TrainPipelineSparseDistOptStash`` / ``EmbStash`` — optimizer state / embedding weights are stashed to pinned host
  memory while the dense part runs and restored before they are needed (memory_stashing).
* ``TrainPipelinePrefetchEMS`` — cache prefetch one batch ahead + embedding-weight stash."""
from __future__ import annotations

from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Callable, Iterator, List, Optional

import torch
from torch import nn

from ..memory_stashing import MemoryStashingManager
from .backward_injection import InjectionSite, register_backward_hook
from .train_pipelines import In, Out, PrefetchTrainPipelineSparseDist, TrainPipelineSparseDist, _to_device


class TrainEvalHybridPipelineBase(TrainPipelineSparseDist[In, Out]):
    def __init__(self, *args: Any, is_eval_batch: Optional[Callable[[Any], bool]] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._is_eval_batch = is_eval_batch or (lambda b: bool(getattr(b, "is_eval", False)))

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self.fill_pipeline(dataloader_iter)
        if not self.batches:
            raise StopIteration
        was_training = self._model.training
        eval_mode = self._is_eval_batch(self.batches[0])
        if eval_mode and was_training:
            self._model.eval()
        try:
            if eval_mode:
                with torch.no_grad():
                    return super().progress(dataloader_iter)
            return super().progress(dataloader_iter)
        finally:
            if eval_mode and was_training:
                self._model.train()


class TrainPipelineSparseDistT(TrainPipelineSparseDist[In, Out]):
    """H2D copies run on a worker thread (the copy stream is still the memcpy stream)."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._copy_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="trb-h2d")
        self._pending_copy: Optional[Future] = None

    def copy_batch_to_gpu(self, dataloader_iter: Iterator[In]):
        batch = self._next_batch(dataloader_iter) if hasattr(self, "_next_batch") else next(dataloader_iter, None)
        if batch is None:
            return None, None
        context = self._create_context()

        def work() -> In:
            with self._stream_context(self._memcpy_stream):
                return _to_device(batch, self._device, non_blocking=True)

        fut = self._copy_pool.submit(work)
        return fut.result(), context


class TrainPipelineSparseDistBwdOpt(TrainPipelineSparseDist[In, Out]):
    """Runs ``injected_work(pipeline)`` from inside the backward pass, when the gradient reaches ``site``."""

    def __init__(self, *args: Any, site: Optional[InjectionSite] = None, injected_work: Optional[Callable[[Any], None]] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._site, self._work, self._handle = site, injected_work, None
        self.injected_calls = 0

    def _ensure_hook(self) -> None:
        if self._handle is None and self._site is not None and self._work is not None:
            inner = self._model.module if hasattr(self._model, "module") else self._model

            def fire(_grad: torch.Tensor) -> None:
                self.injected_calls += 1
                self._work(self)

            self._handle = register_backward_hook(self._site, inner, fire)

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self._ensure_hook()
        return super().progress(dataloader_iter)


class TrainPipelineSparseDistOptStash(TrainPipelineSparseDistBwdOpt[In, Out]):
    """Optimizer state lives on the host during the forward: it is restored (async H2D) when the backward reaches
    ``site`` — just before the fused embedding backward needs it — and stashed again after the optimizer step."""

    def __init__(self, *args: Any, site: Optional[InjectionSite] = None, **kwargs: Any) -> None:
        super().__init__(*args, site=site, injected_work=lambda p: MemoryStashingManager.restore_optimizer_state(), **kwargs)

    def _backward_head(self, losses, head_ctx) -> None:
        if self._site is None:  # no injection site given: the state comes back right before the backward pass starts
            MemoryStashingManager.restore_optimizer_state()
        super()._backward_head(losses, head_ctx)

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        try:
            out = super().progress(dataloader_iter)
        finally:
            MemoryStashingManager.restore_optimizer_state()  # no-op when the hook already restored it
        if self._model.training:
            MemoryStashingManager.stash_optimizer_state(self._model)
        return out


class TrainPipelineSparseDistEmbStash(TrainPipelineSparseDist[In, Out]):
    """Embedding weights are stashed after the embedding forward+backward of a step and restored before the next
    lookup; intended for models whose dense phase is the memory high-water mark."""

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        MemoryStashingManager.restore_embedding_weights()
        out = super().progress(dataloader_iter)
        if self._model.training:
            MemoryStashingManager.stash_embedding_weights(self._model)
        return out

    def detach(self) -> nn.Module:
        MemoryStashingManager.restore_embedding_weights()
        return super().detach()


class TrainPipelinePrefetchEMS(PrefetchTrainPipelineSparseDist[In, Out]):
    """Cache-prefetch pipeline + embedding-memory stash (reference experimental_pipelines.py:1221): the UVM-cache prefetch of batch
    i+1 runs one step ahead as in ``PrefetchTrainPipelineSparseDist`` while the embedding weights of HBM-resident tables are parked
    on the host for the dense phase and come back (async H2D on the stash stream) before the next lookup needs them. Cached tables
    are not stashed - their HBM footprint is the cache the prefetch is filling."""

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        MemoryStashingManager.restore_embedding_weights()
        out = super().progress(dataloader_iter)
        if self._model.training:
            MemoryStashingManager.stash_embedding_weights(self._model)
        return out

    def detach(self) -> nn.Module:
        MemoryStashingManager.restore_embedding_weights()
        return super().detach()


class EvalPipelineCPUSparse:
    """Eval loop for HybridEvalDMP: ``sparse_forward(batch) -> pooled embeddings (CPU)`` runs one batch ahead of
    ``dense_forward(dense_inputs_on_gpu, pooled_on_gpu)``; pooled embeddings travel through reusable pinned buffers."""

    def __init__(self, sparse_forward: Callable[[Any], torch.Tensor], dense_forward: Callable[[Any, torch.Tensor], Any], device: torch.device) -> None:
        self._sparse, self._dense, self._device = sparse_forward, dense_forward, device
        self._pinned: dict = {}
        self._stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self._staged: Optional[tuple] = None

    def _get_or_alloc_pinned(self, name: str, t: torch.Tensor) -> torch.Tensor:
        buf = self._pinned.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=torch.cuda.is_available())
            self._pinned[name] = buf
        buf.copy_(t)
        return buf

    def _stage(self, batch: Any) -> tuple:
        with torch.no_grad():
            pooled = self._sparse(batch)
        pin = self._get_or_alloc_pinned("pooled", pooled)
        if self._stream is not None:
            with torch.cuda.stream(self._stream):
                return batch.to(self._device, non_blocking=True), pin.to(self._device, non_blocking=True)
        return batch, pin.clone()

    def progress(self, dataloader_iter: Iterator[Any]) -> Any:
        if self._staged is None:
            first = next(dataloader_iter, None)
            if first is None:
                raise StopIteration
            self._staged = self._stage(first)
        cur = self._staged
        if self._stream is not None:
            torch.cuda.current_stream(self._device).wait_stream(self._stream)
        nxt = next(dataloader_iter, None)
        self._staged = self._stage(nxt) if nxt is not None else None
        with torch.no_grad():
            out = self._dense(cur[0], cur[1])
        if self._staged is None:
            self._exhausted = True
        return out
