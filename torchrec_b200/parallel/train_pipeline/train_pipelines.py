"""Pipelined training loops (reference torchrec/distributed/train_pipeline/train_pipelines.py).

``TrainPipelineBase``        : H2D copy of batch i+1 (memcpy stream)  ||  fwd/bwd/opt of batch i.
``TrainPipelineSparseDist``  : H2D of batch i+2  ||  sparse input dist of batch i+1 (data_dist stream)
                               ||  fwd/bwd/opt of batch i (default stream).
``TrainPipelineSemiSync``, ``TrainPipelineFusedSparseDist``, ``PrefetchTrainPipelineSparseDist``,
``EvalPipelineSparseDist``, ``StagedTrainPipeline`` build on the same machinery.

Module discovery does not need fx tracing: every ``ShardedModule`` of the model is matched to the
KJT field of the batch that carries its features (falling back to identity recording during the
first forward), and its ``forward`` is swapped for a ``PipelinedForward`` that consumes the
pre-distributed input.
"""
from __future__ import annotations

import abc
from dataclasses import dataclass
import contextlib
import dataclasses
import logging
from collections import deque
from typing import Any, Callable, Deque, Dict, Generic, Iterator, List, Optional, Tuple, Type, TypeVar, Union

import torch
from torch import nn
from torch.autograd.profiler import record_function

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ...streamable import Multistreamable, Pipelineable
from ..model_parallel import DistributedModelParallel
from ..types import Awaitable, ShardedModule
from .pipeline_context import EmbeddingTrainPipelineContext, PrefetchTrainPipelineContext, TrainPipelineContext
from .pipeline_stage import PipelineStage  # noqa: F401  (defined there, its reference import path)

logger = logging.getLogger(__name__)

In = TypeVar("In", bound=Pipelineable)
Out = TypeVar("Out")


def _to_device(batch: In, device: torch.device, non_blocking: bool) -> In:
    assert isinstance(batch, (torch.Tensor, Pipelineable)), f"{type(batch)} must implement Pipelineable interface"
    return batch.to(device=device, non_blocking=non_blocking)


def _wait_for_batch(batch: In, stream: Optional[torch.Stream]) -> None:
    """Make the current stream wait for ``stream`` and tell the caching allocator that ``batch`` is
    now used on the current stream (otherwise its memory could be reused while still read)."""
    if stream is None:
        return
    device = stream.device
    torch.get_device_module(device).current_stream().wait_stream(stream)
    cur_stream = torch.get_device_module(device).current_stream()
    assert isinstance(batch, (torch.Tensor, Multistreamable)), f"{type(batch)} must implement Multistreamable interface"
    batch.record_stream(cur_stream)


def _reduce_losses(losses: torch.Tensor) -> torch.Tensor:
    """Scalar to differentiate: the sum over the leading dim for per-task / per-sample losses, the loss itself when the model
    already returns a scalar (saves a reduction kernel + its autograd node on every step)."""
    return losses if losses.dim() == 0 else torch.sum(losses, dim=0)


def _wait_for_events(batch: In, context: TrainPipelineContext, stream: Optional[torch.Stream]) -> None:
    for event in context.events:
        event.wait()
    context.events.clear()
    if stream is not None:
        batch.record_stream(stream)


class ModelDetachedException(Exception):
    """A pipeline step was asked for while the model is detached (its sharded modules run their plain forward)."""


@dataclass
class TorchCompileConfig:
    """``torch.compile`` options of ``TrainPipelinePT2`` when the caller asks for a tracing compiler (this framework's hot path is
    hand-written kernels + CUDA graphs; nothing compiles unless ``compile_fn`` / this config is given): ``fullgraph`` - one graph or
    error; ``dynamic`` - dynamic shapes (None: automatic); ``backend``; ``compile_on_iter`` - compile at this step, so the first steps
    can be profiled uncompiled."""

    fullgraph: bool = False
    dynamic: Optional[bool] = None
    backend: str = "inductor"
    compile_on_iter: int = 3


class TrainPipeline(abc.ABC, Generic[In, Out]):
    @abc.abstractmethod
    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        ...

    def sync_forward(self) -> None:
        pass

    def reset(self) -> None:
        pass


class TrainPipelineBase(TrainPipeline[In, Out]):
    """Two-stage pipeline: overlap the host-to-device copy of the next batch with the current step."""

    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, device: torch.device,
                 custom_model_fwd: Optional[Callable[[In], Tuple[torch.Tensor, Out]]] = None, enable_inplace_copy_batch: bool = False) -> None:
        self._enable_inplace_copy_batch = enable_inplace_copy_batch  # accepted for parity: batches are moved with one ``to(device, non_blocking)`` per tensor
        self._model = model
        self._optimizer = optimizer
        self._device = device
        self._memcpy_stream: Optional[torch.Stream] = torch.get_device_module(device).Stream() if device.type == "cuda" else None
        self._stream_context = torch.get_device_module(device).stream if device.type == "cuda" else (lambda s: contextlib.nullcontext())
        self._cur_batch: Optional[In] = None
        self._connected = False
        self._data_iter_stopped = False
        self._model_fwd = custom_model_fwd if custom_model_fwd is not None else model

    def _reset_data_iter(self) -> None:
        self._connected = False
        self._data_iter_stopped = False
        self._cur_batch = None

    def _connect(self, dataloader_iter: Iterator[In]) -> None:
        cur_batch = next(dataloader_iter)
        self._cur_batch = cur_batch
        with self._stream_context(self._memcpy_stream):
            self._cur_batch = _to_device(cur_batch, self._device, non_blocking=True)
        self._connected = True

    def _next_batch(self, dataloader_iter: Iterator[In]) -> Optional[In]:
        with record_function("## next_batch ##"):
            try:
                next_batch = next(dataloader_iter)
            except StopIteration:
                self._data_iter_stopped = True
                return None
        return next_batch

    def _wait_for_batch(self, cur_batch: In) -> None:
        with record_function("## wait_for_batch ##"):
            _wait_for_batch(cur_batch, self._memcpy_stream)

    def _backward(self, losses: torch.Tensor) -> None:
        with record_function("## backward ##"):
            _reduce_losses(losses).backward()

    def _copy_batch_to_gpu(self, cur_batch: In) -> None:
        with record_function("## copy_batch_to_gpu ##"):
            with self._stream_context(self._memcpy_stream):
                self._cur_batch = _to_device(cur_batch, self._device, non_blocking=True)

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        if not self._connected:
            self._connect(dataloader_iter)
        if self._data_iter_stopped:
            raise StopIteration()
        cur_batch = self._cur_batch
        assert cur_batch is not None
        if self._model.training:
            with record_function("## zero_grad ##"):
                self._optimizer.zero_grad()
        self._wait_for_batch(cur_batch)
        next_batch = self._next_batch(dataloader_iter)
        if next_batch is not None:
            self._copy_batch_to_gpu(next_batch)
        with record_function("## forward ##"):
            (losses, output) = self._model_fwd(cur_batch)
        if self._model.training:
            self._backward(losses)
            with record_function("## optimizer ##"):
                self._optimizer.step()
        return output


# ---- sparse-dist machinery ----------------------------------------------------------------------------------
class KJTGetter:
    """How to fetch a sharded module's input from a batch: a path of attribute / key accesses."""

    def __init__(self, path: List[Tuple[str, Any]]) -> None:
        self.path = path

    def __call__(self, batch: Any) -> Any:
        obj = batch
        for kind, key in self.path:
            obj = getattr(obj, key) if kind == "attr" else obj[key]
        return obj

    def __repr__(self) -> str:
        return "batch" + "".join(f".{k}" if kind == "attr" else f"[{k!r}]" for kind, k in self.path)


def _iter_kjts(obj: Any, path: List[Tuple[str, Any]], depth: int = 0) -> Iterator[Tuple[List[Tuple[str, Any]], KeyedJaggedTensor]]:
    if isinstance(obj, KeyedJaggedTensor):
        yield path, obj
        return
    if depth > 3 or isinstance(obj, torch.Tensor):
        return
    if dataclasses.is_dataclass(obj):
        for f in dataclasses.fields(obj):
            yield from _iter_kjts(getattr(obj, f.name), path + [("attr", f.name)], depth + 1)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            yield from _iter_kjts(v, path + [("key", k)], depth + 1)
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            yield from _iter_kjts(v, path + [("key", i)], depth + 1)
    elif hasattr(obj, "__dict__") and not isinstance(obj, nn.Module):
        for k, v in vars(obj).items():
            if not k.startswith("_"):
                yield from _iter_kjts(v, path + [("attr", k)], depth + 1)


def _module_feature_names(m: nn.Module) -> Optional[List[str]]:
    for attr in ("_feature_names", "_input_feature_names"):
        names = getattr(m, attr, None)
        if names:
            return list(names)
    return None


class PipelinedForward:
    """Replacement ``forward`` of a pipelined ShardedModule: consume the pre-distributed input of the
    current batch context and run ``compute_and_output_dist`` (reference runtime_forwards.py:92-136)."""

    def __init__(self, name: str, getter: Optional[KJTGetter], module: ShardedModule, pipeline: "TrainPipelineSparseDist", original_forward) -> None:
        self._name = name
        self._getter = getter
        self._module = module
        self._pipeline = pipeline
        self._original_forward = original_forward

    @property
    def name(self) -> str:
        return self._name

    @property
    def getter(self) -> Optional[KJTGetter]:
        return self._getter

    def __call__(self, *input, **kwargs):
        ctx: Optional[TrainPipelineContext] = self._pipeline._context
        if ctx is None or self._name not in ctx.input_dist_tensors_requests:
            if self._getter is None and self._pipeline._recording is not None and input:
                self._pipeline._record_input(self._name, input[0])
            return self._original_forward(*input, **kwargs)
        request = ctx.input_dist_tensors_requests.pop(self._name)
        module_ctx = ctx.module_contexts.pop(self._name)
        stream = self._pipeline._data_dist_stream
        with record_function("## wait_sparse_data_dist ##"):
            with self._pipeline._stream_context(stream):
                data = request.wait()
        if stream is not None:
            cur = torch.get_device_module(self._pipeline._device).current_stream()
            cur.wait_stream(stream)
            data.record_stream(cur)
            module_ctx.record_stream(cur)
        return self._module.compute_and_output_dist(module_ctx, data)


class TrainPipelineSparseDist(TrainPipeline[In, Out]):
    """Three-stage pipeline overlapping H2D copy, sparse input dist and the dense step."""

    _pipelined_forward_type = PipelinedForward

    def __init__(
        self,
        model: nn.Module,
        optimizer: torch.optim.Optimizer,
        device: torch.device,
        execute_all_batches: bool = True,
        apply_jit: bool = False,
        context_type: Type[TrainPipelineContext] = TrainPipelineContext,
        pipeline_postproc: bool = False,
        custom_model_fwd: Optional[Callable[[Optional[In]], Tuple[torch.Tensor, Out]]] = None,
        dmp_collection_sync_interval_batches: Optional[int] = 1,
        enqueue_batch_after_forward: bool = False,
        data_dist_after_forward: bool = False,
        enable_inplace_copy_batch: bool = False,
        free_features_storage_early: bool = False,
        clear_data_dist_inputs: bool = False,
    ) -> None:
        # memory knobs of the reference: ``clear_data_dist_inputs`` / ``free_features_storage_early`` drop the pipeline's references to a
        # batch's input-dist bookkeeping as soon as the collectives have consumed it (before the forward instead of at dequeue)
        self._enable_inplace_copy_batch = enable_inplace_copy_batch
        self._free_features_storage_early = free_features_storage_early
        self._clear_data_dist_inputs = clear_data_dist_inputs or free_features_storage_early
        # data_dist_after_forward: enqueue the head's forward BEFORE the look-ahead input dist. With a device-side input dist (NVLink
        # plane: no host sync, a handful of launches) the only cost of the look-ahead is host time, and spending it first leaves the
        # compute stream idle at the start of every step; the dist kernels still run beside the forward on their own stream.
        self._data_dist_after_forward = data_dist_after_forward
        self._model = model
        self._optimizer = optimizer
        self._device = device
        self._execute_all_batches = execute_all_batches
        self._context_type = context_type
        self._enqueue_batch_after_forward = enqueue_batch_after_forward
        is_cuda = device.type == "cuda"
        self._memcpy_stream: Optional[torch.Stream] = torch.cuda.Stream(device=device, priority=-1) if is_cuda else None
        self._data_dist_stream: Optional[torch.Stream] = torch.cuda.Stream(device=device, priority=-1) if is_cuda else None
        self._stream_context = torch.cuda.stream if is_cuda else (lambda s: contextlib.nullcontext())
        self._model_fwd = custom_model_fwd if custom_model_fwd is not None else model
        self._pipelined_modules: List[ShardedModule] = []
        self._pipelined_forwards: List[PipelinedForward] = []
        self._original_forwards: List[Callable] = []
        self._model_attached = True
        self._batch_i: Optional[In] = None
        self._batch_ip1: Optional[In] = None
        self._batch_ip2: Optional[In] = None
        self._context: Optional[TrainPipelineContext] = None
        self._recording: Optional[Dict[str, Any]] = None
        self._recorded_batch: Any = None
        self.batches: Deque[Optional[In]] = deque()
        self.contexts: Deque[TrainPipelineContext] = deque()
        self._dataloader_iter: Optional[Iterator[In]] = None
        self._dataloader_exhausted: bool = False
        self._initialized = False
        self._dmp_collection_sync_interval_batches = dmp_collection_sync_interval_batches
        self._batch_count = 0
        self._dmp_collection = None
        inner = model.module if isinstance(model, DistributedModelParallel) else model
        if type(inner).__name__ == "DMPCollection" or type(model).__name__ == "DMPCollection":
            self._dmp_collection = model

    # ---- pipeline bookkeeping ------------------------------------------------------------------------------
    def detach(self) -> nn.Module:
        """Restore the original forwards (pipeline can be re-attached)."""
        if self._model_attached:
            for m, fwd in zip(self._pipelined_modules, self._original_forwards):
                m.forward = fwd  # type: ignore[method-assign]
            self._model_attached = False
        return self._model

    def attach(self, model: Optional[nn.Module] = None, sparse_dist: bool = True) -> None:
        if model is not None:
            self._model = model
        self._model_attached = True
        if self._pipelined_modules:
            for m, pf in zip(self._pipelined_modules, self._pipelined_forwards):
                m.forward = pf  # type: ignore[method-assign]

    def _set_module_context(self, context: TrainPipelineContext) -> None:
        self._context = context

    def _record_input(self, name: str, obj: Any) -> None:
        if self._recording is not None:
            self._recording[name] = obj

    def _find_getter(self, batch: Any, module: ShardedModule) -> Optional[KJTGetter]:
        feats = _module_feature_names(module)
        cands = list(_iter_kjts(batch, []))
        if isinstance(batch, KeyedJaggedTensor):
            cands = [([], batch)]
        weighted = bool(getattr(module, "_is_weighted", False))
        match = []
        for path, kjt in cands:
            if feats is not None and not set(feats).issubset(set(kjt.keys())):
                continue
            match.append((path, kjt))
        if len(match) > 1:
            by_w = [(p, k) for p, k in match if (k.weights_or_none() is not None) == weighted]
            if len(by_w) == 1:
                match = by_w
        if len(match) == 1:
            return KJTGetter(match[0][0])
        return None

    def _init_pipelined_modules(self, batch: In, context: TrainPipelineContext) -> None:
        """Find the sharded modules whose input is a KJT carried by the batch and swap their forward."""
        if self._pipelined_modules:
            return
        inner = self._model.module if isinstance(self._model, DistributedModelParallel) else self._model
        for name, m in inner.named_modules():
            if isinstance(m, ShardedModule):
                getter = self._find_getter(batch, m)
                if getter is None:
                    logger.info(f"sharded module {name} is not pipelined (input not found on the batch)")
                pf = self._pipelined_forward_type(name, getter, m, self, m.forward)
                self._pipelined_modules.append(m)
                self._pipelined_forwards.append(pf)
                self._original_forwards.append(m.forward)
                m.forward = pf  # type: ignore[method-assign]
        logger.info(f"pipelined modules: {[(pf.name, pf.getter) for pf in self._pipelined_forwards]}")

    # ---- stages ------------------------------------------------------------------------------------------------
    def _create_context(self) -> TrainPipelineContext:
        context = self._context_type(index=self._batch_count, version=1)
        self._batch_count += 1
        return context

    def _next_batch(self, dataloader_iter: Iterator[In]) -> Optional[In]:
        if dataloader_iter is not self._dataloader_iter:
            self._dataloader_iter = dataloader_iter
            self._dataloader_exhausted = False
        if self._dataloader_exhausted:
            return None
        with record_function("## next_batch ##"):
            batch = next(dataloader_iter, None)
        if batch is None:
            self._dataloader_exhausted = True
        return batch

    def copy_batch_to_gpu(self, dataloader_iter: Iterator[In]) -> Tuple[Optional[In], Optional[TrainPipelineContext]]:
        context = self._create_context()
        with record_function(f"## copy_batch_to_gpu {context.index} ##"):
            with self._stream_context(self._memcpy_stream):
                batch = self._next_batch(dataloader_iter)
                if batch is not None:
                    batch = _to_device(batch, self._device, non_blocking=True)
                elif not self._execute_all_batches:
                    raise StopIteration
                return batch, context

    def enqueue_batch(self, dataloader_iter: Iterator[In]) -> bool:
        batch, context = self.copy_batch_to_gpu(dataloader_iter)
        if batch is None:
            return False
        self.batches.append(batch)
        self.contexts.append(context)
        return True

    def dequeue_batch(self) -> None:
        self.batches.popleft()
        self.contexts.popleft()
        if self.contexts:
            self._set_module_context(self.contexts[0])

    def start_sparse_data_dist(self, batch: Optional[In], context: TrainPipelineContext) -> None:
        """Launch the input dist (splits exchange) of ``batch`` on the data_dist stream."""
        if batch is None:
            return
        with record_function(f"## start_sparse_data_dist {context.index} ##"):
            with self._stream_context(self._data_dist_stream):
                _wait_for_batch(batch, self._memcpy_stream)
                for pf, m in zip(self._pipelined_forwards, self._pipelined_modules):
                    if pf.getter is None:
                        continue
                    kjt = pf.getter(batch)
                    mctx = m.create_context()
                    context.module_contexts[pf.name] = mctx
                    context.input_dist_splits_requests[pf.name] = m.input_dist(mctx, kjt)

    def wait_sparse_data_dist(self, context: TrainPipelineContext) -> None:
        """Finish the splits exchange (host sync) and launch the tensor all-to-alls."""
        with record_function(f"## wait_sparse_data_dist {context.index} ##"):
            with self._stream_context(self._data_dist_stream):
                for name, aw in context.input_dist_splits_requests.items():
                    context.input_dist_tensors_requests[name] = aw.wait()
                context.input_dist_splits_requests.clear()

    def clear_sparse_data_dist_inputs(self, context: TrainPipelineContext) -> None:
        """Drop what the context still holds of a batch's input dist once its collectives are in flight / consumed: the splits
        awaitables (they keep the pre-dist KJT alive)."""
        context.fused_splits_awaitables.clear()
        context.input_dist_splits_requests.clear()

    def fill_pipeline(self, dataloader_iter: Iterator[In]) -> None:
        """Prime the two-deep queue: batch 0 copied + input dist started and finished, batch 1 copied. A queue that already holds
        two batches is full (steady state); with ``execute_all_batches`` the tail of the stream drains a shorter queue."""
        if len(self.batches) >= 2 or (self.batches and self._execute_all_batches):
            return
        primed = bool(self.batches)
        if not primed:
            if not self.enqueue_batch(dataloader_iter):
                return
            self._init_pipelined_modules(self.batches[0], self.contexts[0])
            self.start_sparse_data_dist(self.batches[0], self.contexts[0])
            self.wait_sparse_data_dist(self.contexts[0])
        self.enqueue_batch(dataloader_iter)

    def _late_bind_getters(self, batch: Any) -> None:
        """Resolve modules whose input could only be identified by running a forward."""
        if not self._recording:
            return
        kjts = list(_iter_kjts(batch, []))
        for pf in self._pipelined_forwards:
            if pf.getter is None and pf.name in self._recording:
                obj = self._recording[pf.name]
                for path, kjt in kjts:
                    if kjt is obj:
                        pf._getter = KJTGetter(path)
        self._recording = None

    # One step = the head of the queue trains while the look-ahead work for the next two batches is slotted around its forward:
    #
    #   begin      bind the head's context, clear gradients, make the compute stream wait for the head's input dist
    #   lookahead  (before forward)  start the input dist of batch 1 on the data-dist stream, copy batch 2 on the memcpy stream
    #   forward    model(head): every pipelined module picks its already distributed input out of the context
    #   lookahead  (after forward)   finish batch 1's input dist so that its transfers overlap the backward
    #   backward   loss.backward (fused sparse optimizers run inside) -> 2D weight sync -> dense optimizer
    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self.fill_pipeline(dataloader_iter)
        if not self.batches:
            raise StopIteration
        head, head_ctx = self.batches[0], self.contexts[0]
        self._begin_step(head, head_ctx)
        if self._clear_data_dist_inputs:
            self.clear_sparse_data_dist_inputs(head_ctx)
        self._lookahead(dataloader_iter, forward_done=False)
        losses, output = self._forward_head(head, head_ctx)
        self._lookahead(dataloader_iter, forward_done=True)
        if self._model.training:
            self._backward_head(losses, head_ctx)
        self.dequeue_batch()
        return output

    def _begin_step(self, head: In, head_ctx: TrainPipelineContext) -> None:
        self._set_module_context(head_ctx)
        if self._model.training:
            with record_function("## zero_grad ##"):
                self._optimizer.zero_grad()
        with record_function("## wait_for_batch ##"):
            _wait_for_batch(head, self._data_dist_stream)

    def _lookahead(self, dataloader_iter: Iterator[In], forward_done: bool) -> None:
        have_next = len(self.batches) >= 2
        late_dist = getattr(self, "_data_dist_after_forward", False)
        if not forward_done:
            if have_next and not late_dist:
                self.start_sparse_data_dist(self.batches[1], self.contexts[1])
            if not self._enqueue_batch_after_forward:
                self.enqueue_batch(dataloader_iter)
            return
        if have_next and late_dist:
            self.start_sparse_data_dist(self.batches[1], self.contexts[1])
        if self._enqueue_batch_after_forward:
            self.enqueue_batch(dataloader_iter)
        if have_next:
            self.wait_sparse_data_dist(self.contexts[1])

    def _forward_head(self, head: In, head_ctx: TrainPipelineContext):
        unresolved = any(pf.getter is None for pf in self._pipelined_forwards)
        if unresolved and self._recording is None and not self._initialized:
            self._recording = {}
        with record_function(f"## forward {head_ctx.index} ##"):
            losses, output = self._model_fwd(head)
        if self._recording is not None:
            self._late_bind_getters(head)
        self._initialized = True
        return losses, output

    def _backward_head(self, losses, head_ctx: TrainPipelineContext) -> None:
        with record_function(f"## backward {head_ctx.index} ##"):
            _reduce_losses(losses).backward()
        self.sync_embeddings()
        with record_function(f"## optimizer {head_ctx.index} ##"):
            self._optimizer.step()

    def sync_embeddings(self) -> None:
        """2D-parallel weight sync across replica groups every N batches (reference :208-239)."""
        m = self._dmp_collection
        interval = self._dmp_collection_sync_interval_batches
        if m is None or not interval:
            return
        if self.contexts and self.contexts[0].index is not None and self.contexts[0].index % interval == 0:
            m.sync()

    def reset(self) -> None:
        self.batches.clear()
        self.contexts.clear()
        self._context = None
        self._dataloader_iter = None
        self._dataloader_exhausted = False


class TrainPipelineSparseDistLite(TrainPipelineSparseDist[In, Out]):
    """Memory-lean variant: only one extra batch in flight; the input dist of batch i+1 starts
    after the forward of batch i (overlaps with backward only). Reference :1259-1435."""

    def fill_pipeline(self, dataloader_iter: Iterator[In]) -> None:
        if self.batches:
            return
        if not self.enqueue_batch(dataloader_iter):
            return
        self._init_pipelined_modules(self.batches[0], self.contexts[0])
        self.start_sparse_data_dist(self.batches[0], self.contexts[0])
        self.wait_sparse_data_dist(self.contexts[0])

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self.fill_pipeline(dataloader_iter)
        if not self.batches:
            raise StopIteration
        self._set_module_context(self.contexts[0])
        if self._model.training:
            self._optimizer.zero_grad()
        _wait_for_batch(self.batches[0], self._data_dist_stream)
        losses, output = self._model_fwd(self.batches[0])
        has_next = self.enqueue_batch(dataloader_iter)
        if has_next:
            self.start_sparse_data_dist(self.batches[1], self.contexts[1])
        if self._model.training:
            _reduce_losses(losses).backward()
        if has_next:
            self.wait_sparse_data_dist(self.contexts[1])
        if self._model.training:
            self._optimizer.step()
        self.dequeue_batch()
        return output


class EvalPipelineSparseDist(TrainPipelineSparseDist[In, Out]):
    """Pipelined evaluation (no backward / optimizer). Reference :2269-2410."""

    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, device: torch.device, apply_jit: bool = False, pipeline_postproc: bool = False,
                 **memory_knobs: Any) -> None:
        super().__init__(model, optimizer, device, True, apply_jit, pipeline_postproc=pipeline_postproc, **memory_knobs)

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self.fill_pipeline(dataloader_iter)
        if not self.batches:
            raise StopIteration
        self._set_module_context(self.contexts[0])
        _wait_for_batch(self.batches[0], self._data_dist_stream)
        if len(self.batches) >= 2:
            self.start_sparse_data_dist(self.batches[1], self.contexts[1])
        self.enqueue_batch(dataloader_iter)
        with torch.no_grad():
            losses, output = self._model_fwd(self.batches[0])
        if len(self.batches) >= 2:
            self.wait_sparse_data_dist(self.contexts[1])
        self.dequeue_batch()
        return output


class TrainPipelineSemiSync(TrainPipelineSparseDist[In, Out]):
    """Semi-synchronous training: the embedding lookup + output dist of batch B+1 is launched
    *before* the optimizer step of batch B (embeddings are stale by one step), overlapping the
    embedding all-to-all with the dense forward of batch B. Reference :1650-1976.

    ``start_batch`` selects the first batch that runs semi-synchronously."""

    def __init__(self, model, optimizer, device, execute_all_batches: bool = True, apply_jit: bool = False, start_batch: int = 900,
                 stash_gradients: bool = False, pipeline_postproc: bool = True, custom_model_fwd=None, strict: bool = False,
                 dmp_collection_sync_interval_batches: Optional[int] = 1, **memory_knobs: Any) -> None:
        super().__init__(model, optimizer, device, execute_all_batches, apply_jit, context_type=EmbeddingTrainPipelineContext,
                         pipeline_postproc=pipeline_postproc, custom_model_fwd=custom_model_fwd, dmp_collection_sync_interval_batches=dmp_collection_sync_interval_batches,
                         **memory_knobs)
        self._start_batch = start_batch
        self._stash_gradients = stash_gradients
        self._embedding_streams_enabled = device.type == "cuda"
        self._embedding_stream: Optional[torch.Stream] = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self._precomputed: Dict[int, Dict[str, Any]] = {}

    def is_semi_sync(self) -> bool:
        return bool(self.contexts) and self.contexts[0].index is not None and self.contexts[0].index >= self._start_batch

    def _start_embedding_lookup(self, batch: In, context: TrainPipelineContext) -> None:
        """Run compute_and_output_dist of every pipelined module for ``batch`` now (stale weights)."""
        outs: Dict[str, Any] = {}
        with self._stream_context(self._embedding_stream):
            if self._embedding_stream is not None:
                self._embedding_stream.wait_stream(self._data_dist_stream)
            for pf, m in zip(self._pipelined_forwards, self._pipelined_modules):
                if pf.name in context.input_dist_tensors_requests:
                    data = context.input_dist_tensors_requests.pop(pf.name).wait()
                    mctx = context.module_contexts.pop(pf.name)
                    outs[pf.name] = m.compute_and_output_dist(mctx, data)
        self._precomputed[id(context)] = outs

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self.fill_pipeline(dataloader_iter)
        if not self.batches:
            raise StopIteration
        ctx0 = self.contexts[0]
        self._set_module_context(ctx0)
        if self._model.training:
            self._optimizer.zero_grad()
        _wait_for_batch(self.batches[0], self._data_dist_stream)
        if len(self.batches) >= 2:
            self.start_sparse_data_dist(self.batches[1], self.contexts[1])
        self.enqueue_batch(dataloader_iter)
        pre = self._precomputed.pop(id(ctx0), None)
        if pre is not None:
            # embeddings were computed ahead of time: feed them through patched forwards
            saved = []
            for pf, m in zip(self._pipelined_forwards, self._pipelined_modules):
                if pf.name in pre:
                    res = pre[pf.name]
                    saved.append((m, m.forward))
                    m.forward = (lambda r: (lambda *a, **k: r))(res)  # type: ignore[method-assign]
            if self._embedding_stream is not None:
                torch.cuda.current_stream().wait_stream(self._embedding_stream)
            try:
                losses, output = self._model_fwd(self.batches[0])
            finally:
                for m, f in saved:
                    m.forward = f  # type: ignore[method-assign]
        else:
            losses, output = self._model_fwd(self.batches[0])
        if len(self.batches) >= 2:
            self.wait_sparse_data_dist(self.contexts[1])
            if self.contexts[1].index is not None and self.contexts[1].index >= self._start_batch:
                self._start_embedding_lookup(self.batches[1], self.contexts[1])
        if self._model.training:
            _reduce_losses(losses).backward()
            self._optimizer.step()
        self.dequeue_batch()
        return output


class TrainPipelineFusedSparseDist(TrainPipelineSparseDist[In, Out]):
    """Also runs the embedding lookup of batch i+1 on a separate stream so it overlaps with the
    optimizer of batch i (reference :1437-1648). The lookup reads weights before the dense optimizer
    step but after the (in-backward) sparse update, so results are identical to the base pipeline."""

    def __init__(self, model, optimizer, device, execute_all_batches: bool = True, apply_jit: bool = False, pipeline_postproc: bool = True,
                 custom_model_fwd=None, embedding_lookup_after_opt: bool = False, strict: bool = False, emb_lookup_stream: str = "data_dist") -> None:
        super().__init__(model, optimizer, device, execute_all_batches, apply_jit, context_type=EmbeddingTrainPipelineContext,
                         pipeline_postproc=pipeline_postproc, custom_model_fwd=custom_model_fwd)
        self._embedding_lookup_after_opt = embedding_lookup_after_opt
        if emb_lookup_stream == "new" and device.type == "cuda":
            self._emb_stream: Optional[torch.Stream] = torch.cuda.Stream(device=device)
        elif emb_lookup_stream == "current" or device.type != "cuda":
            self._emb_stream = None
        else:
            self._emb_stream = self._data_dist_stream
        self._precomputed: Dict[int, Dict[str, Any]] = {}

    def _embedding_lookup(self, context: TrainPipelineContext) -> None:
        outs: Dict[str, Any] = {}
        with self._stream_context(self._emb_stream):
            for pf, m in zip(self._pipelined_forwards, self._pipelined_modules):
                if pf.name in context.input_dist_tensors_requests:
                    data = context.input_dist_tensors_requests.pop(pf.name).wait()
                    mctx = context.module_contexts.pop(pf.name)
                    outs[pf.name] = m.compute_and_output_dist(mctx, data)
        self._precomputed[id(context)] = outs

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self.fill_pipeline(dataloader_iter)
        if not self.batches:
            raise StopIteration
        ctx0 = self.contexts[0]
        self._set_module_context(ctx0)
        if self._model.training:
            self._optimizer.zero_grad()
        _wait_for_batch(self.batches[0], self._data_dist_stream)
        if len(self.batches) >= 2:
            self.start_sparse_data_dist(self.batches[1], self.contexts[1])
        self.enqueue_batch(dataloader_iter)
        pre = self._precomputed.pop(id(ctx0), None)
        saved = []
        if pre:
            if self._emb_stream is not None:
                torch.cuda.current_stream().wait_stream(self._emb_stream)
            for pf, m in zip(self._pipelined_forwards, self._pipelined_modules):
                if pf.name in pre:
                    saved.append((m, m.forward))
                    m.forward = (lambda r: (lambda *a, **k: r))(pre[pf.name])  # type: ignore[method-assign]
        try:
            losses, output = self._model_fwd(self.batches[0])
        finally:
            for m, f in saved:
                m.forward = f  # type: ignore[method-assign]
        if len(self.batches) >= 2:
            self.wait_sparse_data_dist(self.contexts[1])
        if self._model.training:
            _reduce_losses(losses).backward()
            if len(self.batches) >= 2 and not self._embedding_lookup_after_opt:
                if self._emb_stream is not None:
                    self._emb_stream.wait_stream(torch.cuda.current_stream())
                self._embedding_lookup(self.contexts[1])
            self._optimizer.step()
            if len(self.batches) >= 2 and self._embedding_lookup_after_opt:
                self._embedding_lookup(self.contexts[1])
        self.dequeue_batch()
        return output


class EvalPipelineFusedSparseDist(TrainPipelineFusedSparseDist[In, Out]):
    """Evaluation through the fused pipeline: the embedding lookup of batch i+1 runs on the lookup stream while batch i is in the dense
    forward; no backward, no optimizer step (in training mode the gradients are only cleared). A detached model is re-attached by
    ``progress``. Reference train_pipelines.py:2396-2461."""

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        if not self._model_attached:
            self.attach(self._model)
        self.fill_pipeline(dataloader_iter)
        if not self.batches:
            raise StopIteration
        ctx0 = self.contexts[0]
        self._set_module_context(ctx0)
        if self._model.training:
            self._optimizer.zero_grad()
        _wait_for_batch(self.batches[0], self._data_dist_stream)
        if len(self.batches) >= 2:
            self.start_sparse_data_dist(self.batches[1], self.contexts[1])
        self.enqueue_batch(dataloader_iter)
        pre = self._precomputed.pop(id(ctx0), None)
        saved = []
        if pre:
            if self._emb_stream is not None:
                torch.cuda.current_stream().wait_stream(self._emb_stream)
            for pf, m in zip(self._pipelined_forwards, self._pipelined_modules):
                if pf.name in pre:
                    saved.append((m, m.forward))
                    m.forward = (lambda r: (lambda *a, **k: r))(pre[pf.name])  # type: ignore[method-assign]
        try:
            with torch.no_grad():
                losses, output = self._model_fwd(self.batches[0])
        finally:
            for m, f in saved:
                m.forward = f  # type: ignore[method-assign]
        if len(self.batches) >= 2:
            self.wait_sparse_data_dist(self.contexts[1])
            if self._emb_stream is not None:
                self._emb_stream.wait_stream(torch.cuda.current_stream())
            self._embedding_lookup(self.contexts[1])
        self.dequeue_batch()
        return output


class PrefetchTrainPipelineSparseDist(TrainPipelineSparseDist[In, Out]):
    """Four-stage pipeline adding a cache-prefetch stage for host-offloaded (UVM-caching style)
    tables: the rows a batch will touch are staged into the HBM cache on a prefetch stream one
    step ahead (reference :1978-2267)."""

    def __init__(self, model, optimizer, device, execute_all_batches: bool = True, apply_jit: bool = False, pipeline_postproc: bool = True,
                 custom_model_fwd=None, **memory_knobs: Any) -> None:
        super().__init__(model, optimizer, device, execute_all_batches, apply_jit, context_type=PrefetchTrainPipelineContext,
                         pipeline_postproc=pipeline_postproc, custom_model_fwd=custom_model_fwd, **memory_knobs)
        self._prefetch_stream: Optional[torch.Stream] = torch.cuda.Stream(device=device) if device.type == "cuda" else None

    def _prefetch(self, context: TrainPipelineContext) -> None:
        with self._stream_context(self._prefetch_stream):
            if self._prefetch_stream is not None:
                self._prefetch_stream.wait_stream(self._data_dist_stream)
            for pf, m in zip(self._pipelined_forwards, self._pipelined_modules):
                if pf.name in context.input_dist_tensors_requests and hasattr(m, "prefetch"):
                    req = context.input_dist_tensors_requests[pf.name]
                    data = req.wait()
                    m.prefetch(context.module_contexts[pf.name], data)
                    from ..types import NoWait

                    context.input_dist_tensors_requests[pf.name] = NoWait(data)

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self.fill_pipeline(dataloader_iter)
        if not self.batches:
            raise StopIteration
        self._set_module_context(self.contexts[0])
        if self._model.training:
            self._optimizer.zero_grad()
        _wait_for_batch(self.batches[0], self._data_dist_stream)
        if self._prefetch_stream is not None:
            torch.cuda.current_stream().wait_stream(self._prefetch_stream)
        if len(self.batches) >= 2:
            self.start_sparse_data_dist(self.batches[1], self.contexts[1])
        self.enqueue_batch(dataloader_iter)
        losses, output = self._model_fwd(self.batches[0])
        if len(self.batches) >= 2:
            self.wait_sparse_data_dist(self.contexts[1])
            self._prefetch(self.contexts[1])
        if self._model.training:
            _reduce_losses(losses).backward()
            self._optimizer.step()
        self.dequeue_batch()
        return output




class StagedTrainPipeline(TrainPipeline[In, Optional[In]]):
    """Generic software pipeline over user stages: ``progress`` returns a batch that went through
    all stages while later batches advance one stage each (reference :2412-2700)."""

    def __init__(self, pipeline_stages: List[PipelineStage], debug_mode: bool = False, compute_stream: Optional[torch.Stream] = None,
                 on_flush_end: Optional[Callable[[], None]] = None) -> None:
        self._pipeline_stages = pipeline_stages
        self._debug_mode = debug_mode
        self._stage_outputs: List[Optional[Tuple[Any, Optional[torch.Event]]]] = [None] * len(pipeline_stages)
        self._initialized = False
        self._num_steps = 0
        self._dataloader_iter: Optional[Iterator[In]] = None
        self._dataloader_exhausted = False
        self._compute_stream = compute_stream or (torch.cuda.current_stream() if torch.cuda.is_available() else None)
        self._flushing = False
        self.on_flush_end = on_flush_end

    @property
    def num_stages(self) -> int:
        return len(self._pipeline_stages)

    def _next_batch(self, dataloader_iter: Iterator[In]) -> Optional[In]:
        if dataloader_iter is not self._dataloader_iter:
            self._dataloader_iter = dataloader_iter
            self._dataloader_exhausted = False
        if self._dataloader_exhausted or self._flushing:
            return None
        batch = next(dataloader_iter, None)
        if batch is None:
            self._dataloader_exhausted = True
        return batch

    def _run_stage(self, stage_idx: int, batch: Optional[Any]) -> Optional[Tuple[Any, Optional[torch.Event]]]:
        if batch is None:
            return None
        stage = self._pipeline_stages[stage_idx]
        ctx = torch.cuda.stream(stage.stream) if stage.stream is not None else contextlib.nullcontext()
        with ctx:
            out = stage.runnable(batch)
            ev = None
            if stage.stream is not None:
                ev = torch.cuda.Event()
                ev.record(stage.stream)
        return out, ev

    def progress(self, dataloader_iter: Iterator[In]) -> Optional[In]:
        n = self.num_stages
        if not self._initialized:
            # fill: batch k goes through stages 0..n-1-k
            for k in range(n - 1):
                b: Any = self._next_batch(dataloader_iter)
                res: Optional[Tuple[Any, Optional[torch.Event]]] = (b, None) if b is not None else None
                for s in range(0, n - 1 - k):
                    if res is None:
                        break
                    if res[1] is not None and self._pipeline_stages[s].stream is not None:
                        self._pipeline_stages[s].stream.wait_event(res[1])
                    res = self._run_stage(s, res[0])
                self._stage_outputs[n - 2 - k] = res
            for st in self._pipeline_stages:
                if st.fill_callback is not None:
                    st.fill_callback()
            self._initialized = True
        # advance: oldest batch leaves through the last stage
        new_outputs: List[Optional[Tuple[Any, Optional[torch.Event]]]] = [None] * n
        for s in range(n - 1, -1, -1):
            inp = self._stage_outputs[s - 1] if s > 0 else None
            if s == 0:
                b = self._next_batch(dataloader_iter)
                inp = (b, None) if b is not None else None
            if inp is None:
                continue
            if inp[1] is not None and self._pipeline_stages[s].stream is not None:
                self._pipeline_stages[s].stream.wait_event(inp[1])
            new_outputs[s] = self._run_stage(s, inp[0])
        out = new_outputs[n - 1]
        self._stage_outputs = new_outputs
        self._num_steps += 1
        if out is None:
            if all(o is None for o in new_outputs):
                if self._flushing:
                    self._flushing = False
                    self._initialized = False
                    if self.on_flush_end is not None:
                        self.on_flush_end()
                return None
            return self.progress(dataloader_iter) if any(o is not None for o in new_outputs[:-1]) else None
        if out[1] is not None and self._compute_stream is not None:
            self._compute_stream.wait_event(out[1])
        return out[0]

    def set_flush(self, flag: bool) -> None:
        self._flushing = flag

    def flush_end(self) -> None:
        self._flushing = False


class TrainPipelinePT2(TrainPipelineBase[In, Out]):
    """``TrainPipelineBase`` whose model is handed to a graph compiler after ``pre_compile_fn`` has seen a batch (reference
    train_pipelines.py:432-537, which calls ``torch.compile``). This framework's hot path is hand-written kernels + CUDA graphs, not a
    tracing compiler, so the default ``compile_fn`` captures the dense sub-modules into CUDA graphs when the model offers
    ``capture_dense_graphs`` and otherwise leaves the model untouched; pass ``compile_fn=torch.compile`` to get the reference
    behaviour."""

    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, device: torch.device, compile_fn: Optional[Callable[[nn.Module], nn.Module]] = None,
                 pre_compile_fn: Optional[Callable[[nn.Module], None]] = None, post_compile_fn: Optional[Callable[[nn.Module], None]] = None,
                 input_transformer: Optional[Callable[[In], In]] = None, num_pre_compile_steps: int = 1,
                 torch_compile_config: Optional[TorchCompileConfig] = None) -> None:
        super().__init__(model, optimizer, device)
        if torch_compile_config is not None and compile_fn is None:
            cfg = torch_compile_config
            compile_fn = lambda m: torch.compile(m, fullgraph=cfg.fullgraph, dynamic=cfg.dynamic, backend=cfg.backend)  # noqa: E731
            num_pre_compile_steps = cfg.compile_on_iter
        self._compile_fn = compile_fn
        self._pre_compile_fn = pre_compile_fn
        self._post_compile_fn = post_compile_fn
        self._input_transformer = input_transformer
        self._num_pre_compile_steps = max(0, num_pre_compile_steps)
        self._steps = 0
        self._compiled = False

    def _maybe_compile(self) -> None:
        if self._compiled or self._steps < self._num_pre_compile_steps:
            return
        if self._pre_compile_fn is not None:
            self._pre_compile_fn(self._model)
        if self._compile_fn is not None:
            self._model = self._compile_fn(self._model)
        if self._post_compile_fn is not None:
            self._post_compile_fn(self._model)
        self._compiled = True

    def _next_batch(self, dataloader_iter: Iterator[In]) -> Optional[In]:
        batch = super()._next_batch(dataloader_iter)
        if batch is not None and self._input_transformer is not None:
            batch = self._input_transformer(batch)
        return batch

    def _connect(self, dataloader_iter: Iterator[In]) -> None:
        super()._connect(dataloader_iter)
        if self._cur_batch is not None and self._input_transformer is not None:
            self._cur_batch = self._input_transformer(self._cur_batch)

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        self._maybe_compile()
        out = super().progress(dataloader_iter)
        self._steps += 1
        return out


class TrainPipelineSparseDistCompAutograd(TrainPipelineSparseDist[In, Out]):
    """``TrainPipelineSparseDist`` whose backward runs under compiled autograd when the installed torch provides it and
    ``enable`` is set (reference train_pipelines.py:2269-2330). Off by default: the fused backward + optimizer kernels already are
    the backward graph, and compiled autograd cannot trace through the ctypes launches."""

    def __init__(self, *args: Any, enable: bool = False, compiler_fn: Optional[Callable[..., Any]] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._ca_enable = enable
        self._ca_compiler_fn = compiler_fn

    def _backward_ctx(self):
        if not self._ca_enable:
            return contextlib.nullcontext()
        try:
            from torch._dynamo import compiled_autograd

            return compiled_autograd._enable(self._ca_compiler_fn or (lambda gm: gm))
        except Exception:
            return contextlib.nullcontext()

    def progress(self, dataloader_iter: Iterator[In]) -> Out:
        with self._backward_ctx():
            return super().progress(dataloader_iter)
