"""Per-batch pipeline state (reference train_pipeline/pipeline_context.py:27)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch

from ...streamable import Multistreamable
from ..types import Awaitable


@dataclass
class TrainPipelineContext:
    """Carries one batch's in-flight input-dist requests between pipeline stages."""

    input_dist_splits_requests: Dict[str, Awaitable[Any]] = field(default_factory=dict)
    input_dist_tensors_requests: Dict[str, Awaitable[Any]] = field(default_factory=dict)
    module_contexts: Dict[str, Multistreamable] = field(default_factory=dict)
    module_contexts_next_batch: Dict[str, Multistreamable] = field(default_factory=dict)
    fused_splits_awaitables: List[Any] = field(default_factory=list)
    events: List[torch.Event] = field(default_factory=list)
    postproc_fwd_results: Dict[str, Any] = field(default_factory=dict)
    index: Optional[int] = None
    version: int = 1


@dataclass
class PrefetchTrainPipelineContext(TrainPipelineContext):
    module_input_post_prefetch: Dict[str, Multistreamable] = field(default_factory=dict)
    module_contexts_post_prefetch: Dict[str, Multistreamable] = field(default_factory=dict)
    module_input_post_prefetch_next_batch: Dict[str, Multistreamable] = field(default_factory=dict)
    module_contexts_post_prefetch_next_batch: Dict[str, Multistreamable] = field(default_factory=dict)


@dataclass
class EmbeddingTrainPipelineContext(TrainPipelineContext):
    embedding_a2a_requests: Dict[str, Any] = field(default_factory=dict)
    embedding_tensors: List[List[torch.Tensor]] = field(default_factory=list)
    embedding_features: List[List[Any]] = field(default_factory=list)
    detached_embedding_tensors: List[List[torch.Tensor]] = field(default_factory=list)


@dataclass
class CPUEmbeddingTrainPipelineContext(EmbeddingTrainPipelineContext):
    """Embedding lookup on the host (tables in host memory / a parameter server), dense part on a GPU: the stage that copies the looked-up
    embeddings to the device stores them here, keyed by the fqn of the embedding module (reference pipeline_context.py:101)."""

    dense_gpu_device: str = field(default_factory=str)
    gpu_embedding_outputs: Dict[str, Any] = field(default_factory=dict)


In = Any
Out = Any
