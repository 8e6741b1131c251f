"""Train-pipeline selection (reference ``torchrec/distributed/test_utils/pipeline_config.py:40`` ``PipelineConfig``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
from torch import nn

from .. import train_pipeline as tp


@dataclass
class PipelineConfig:
    pipeline: str = "sparse"  # base | sparse | sparse_lite | fused | semi | prefetch
    emb_lookup_stream: str = "data_dist"
    apply_jit: bool = False
    inplace_copy_batch_to_gpu: bool = False

    _MAP = {"base": "TrainPipelineBase", "sparse": "TrainPipelineSparseDist", "sparse_lite": "TrainPipelineSparseDistLite", "fused": "TrainPipelineFusedSparseDist",
            "semi": "TrainPipelineSemiSync", "prefetch": "PrefetchTrainPipelineSparseDist"}

    def generate_pipeline(self, model: nn.Module, opt: torch.optim.Optimizer, device: torch.device, **kwargs: Any):
        if self.pipeline not in self._MAP:
            raise ValueError(f"unknown pipeline {self.pipeline!r}; available: {sorted(self._MAP)}")
        cls = getattr(tp, self._MAP[self.pipeline])
        return cls(model=model, optimizer=opt, device=device, **kwargs)
