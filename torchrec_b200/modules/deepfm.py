"""DeepFM building blocks (reference torchrec/modules/deepfm.py:36-230)."""
from typing import List

import torch
from torch import nn


def _get_flatten_input(inputs: List[torch.Tensor]) -> torch.Tensor:
    return torch.cat([inp.flatten(1) for inp in inputs], dim=1)


class DeepFM(nn.Module):
    """Deep part: flatten + concat all inputs, then a user supplied dense module."""

    def __init__(self, dense_module: nn.Module) -> None:
        super().__init__()
        self.dense_module = dense_module

    def forward(self, embeddings: List[torch.Tensor]) -> torch.Tensor:
        return self.dense_module(_get_flatten_input(embeddings))


class FactorizationMachine(nn.Module):
    """Second-order FM term: 0.5 * ((sum x)^2 - sum x^2) summed over the feature dim."""

    def forward(self, embeddings: List[torch.Tensor]) -> torch.Tensor:
        fm_input = _get_flatten_input(embeddings)
        sum_of_input = torch.sum(fm_input, dim=1, keepdim=True)
        sum_of_square = torch.sum(fm_input * fm_input, dim=1, keepdim=True)
        square_of_sum = sum_of_input * sum_of_input
        return (square_of_sum - sum_of_square) * 0.5
