"""Activation modules (reference torchrec/modules/activation.py:20)."""
from typing import List, Optional, Union

import torch
from torch import nn


class SwishLayerNorm(nn.Module):
    """``x * sigmoid(LayerNorm(x))``."""

    def __init__(self, input_dims: Union[int, List[int], torch.Size], device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.norm: nn.Module = nn.Sequential(nn.LayerNorm(input_dims, device=device), nn.Sigmoid())

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return input * self.norm(input)
