"""Packed sub-byte tensors (reference torchrec/tensor_types.py:170,181): ``UInt4Tensor`` / ``UInt2Tensor`` wrap a uint8
storage holding 2 / 4 elements per byte — the logical view of INT4 / INT2 quantized embedding rows (without the fused
scale/bias tail) so that state-dict tooling can reason about shapes in elements."""
from typing import Any, Tuple

import torch


class _UIntXTensor:
    BITS = 8

    def __init__(self, data: torch.Tensor) -> None:
        assert data.dtype == torch.uint8 and data.dim() == 2, "packed storage must be a 2-D uint8 tensor"
        self.elem = data

    @property
    def per_byte(self) -> int:
        return 8 // self.BITS

    @property
    def shape(self) -> torch.Size:
        return torch.Size([self.elem.shape[0], self.elem.shape[1] * self.per_byte])

    def size(self, dim: int = None):  # type: ignore[assignment]
        return self.shape if dim is None else self.shape[dim]

    def dim(self) -> int:
        return 2

    @property
    def device(self) -> torch.device:
        return self.elem.device

    @property
    def dtype(self) -> str:
        return f"uint{self.BITS}"

    def to(self, *args: Any, **kwargs: Any):
        return type(self)(self.elem.to(*args, **kwargs))

    def detach(self):
        return type(self)(self.elem.detach())

    def clone(self):
        return type(self)(self.elem.clone())

    def view(self, dtype: torch.dtype) -> torch.Tensor:
        assert dtype == torch.uint8, "packed tensors can only be viewed as uint8 bytes"
        return self.elem

    def __getitem__(self, idx):
        rows = self.elem[idx]
        return type(self)(rows if rows.dim() == 2 else rows.unsqueeze(0))

    def unpack(self) -> torch.Tensor:
        """uint8 tensor with one element per entry (low bits first within a byte)."""
        mask = (1 << self.BITS) - 1
        parts = [(self.elem >> (self.BITS * k)) & mask for k in range(self.per_byte)]
        return torch.stack(parts, dim=-1).reshape(self.elem.shape[0], -1)

    @classmethod
    def pack(cls, values: torch.Tensor):
        per = 8 // cls.BITS
        assert values.dim() == 2 and values.shape[1] % per == 0
        v = values.to(torch.uint8).reshape(values.shape[0], -1, per)
        out = torch.zeros(v.shape[:2], dtype=torch.uint8, device=values.device)
        for k in range(per):
            out |= (v[..., k] & ((1 << cls.BITS) - 1)) << (cls.BITS * k)
        return cls(out)

    def __repr__(self) -> str:
        return f"{type(self).__name__}(shape={tuple(self.shape)}, device={self.device})"


class UInt4Tensor(_UIntXTensor):
    BITS = 4


class UInt2Tensor(_UIntXTensor):
    BITS = 2
