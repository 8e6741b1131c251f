"""Batch generation config (reference ``torchrec/distributed/test_utils/input_config.py:20`` ``ModelInputConfig``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Optional

import torch

from .model_input import ModelInput


@dataclass
class ModelInputConfig:
    num_batches: int = 10
    batch_size: int = 8192
    num_float_features: int = 10
    feature_pooling_avg: int = 10
    use_offsets: bool = False
    dev_str: str = ""
    long_kjt_indices: bool = True
    long_kjt_offsets: bool = True
    long_kjt_lengths: bool = True
    pin_memory: bool = True
    power_law_alpha: Optional[float] = None

    def generate_batches(self, tables: List[Any], weighted_tables: List[Any]) -> List[ModelInput]:
        device = torch.device(self.dev_str) if self.dev_str else None
        i64 = lambda b: torch.int64 if b else torch.int32  # noqa: E731
        return [ModelInput.generate(batch_size=self.batch_size, tables=tables, weighted_tables=weighted_tables, num_float_features=self.num_float_features,
                                    pooling_avg=self.feature_pooling_avg, use_offsets=self.use_offsets, device=device, indices_dtype=i64(self.long_kjt_indices),
                                    offsets_dtype=i64(self.long_kjt_offsets), lengths_dtype=i64(self.long_kjt_lengths), pin_memory=self.pin_memory and device is None,
                                    power_law_alpha=self.power_law_alpha) for _ in range(self.num_batches)]
