"""Stream-safety protocols (reference torchrec/streamable.py:15-48)."""
from __future__ import annotations

import abc

import torch


class Multistreamable(abc.ABC):
    """Objects that cross CUDA streams must tell the caching allocator which streams use them."""

    @abc.abstractmethod
    def record_stream(self, stream: torch.Stream) -> None:
        ...


class Pipelineable(Multistreamable):
    """Objects a train pipeline can move to the device asynchronously."""

    @abc.abstractmethod
    def to(self, device: torch.device, non_blocking: bool) -> "Pipelineable":
        ...
