"""A list of tensors handed to the native side as raw pointers (reference ``contrib/dynamic_embedding/.../tensor_list.py:21``).

The parameter-server client and the id transformer take ``TensorList``s: the tensors stay owned by Python (kept alive by this object), the native code
sees ``(data_ptr, nbytes, dtype code)`` triples - no copies, no torch headers needed in the C++ core."""
from __future__ import annotations

import ctypes
from typing import Iterator, List

import torch

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.int64: 3, torch.int32: 4, torch.uint8: 5, torch.float64: 6}


class TensorList:
    def __init__(self, tensors: List[torch.Tensor]) -> None:
        for t in tensors:
            if not t.is_contiguous():
                raise ValueError("TensorList needs contiguous tensors (the native side indexes raw memory)")
            if t.dtype not in _DTYPE_CODE:
                raise ValueError(f"unsupported dtype {t.dtype}")
        self._tensors = list(tensors)

    def __len__(self) -> int:
        return len(self._tensors)

    def __getitem__(self, i: int) -> torch.Tensor:
        return self._tensors[i]

    def __iter__(self) -> Iterator[torch.Tensor]:
        return iter(self._tensors)

    def append(self, t: torch.Tensor) -> None:
        self.__init__(self._tensors + [t])

    def pointers(self) -> "ctypes.Array[ctypes.c_void_p]":
        return (ctypes.c_void_p * len(self._tensors))(*[t.data_ptr() for t in self._tensors])

    def nbytes(self) -> "ctypes.Array[ctypes.c_int64]":
        return (ctypes.c_int64 * len(self._tensors))(*[t.numel() * t.element_size() for t in self._tensors])

    def dtype_codes(self) -> "ctypes.Array[ctypes.c_int32]":
        return (ctypes.c_int32 * len(self._tensors))(*[_DTYPE_CODE[t.dtype] for t in self._tensors])

    def row_bytes(self) -> List[int]:
        """Bytes of one row (dim 0 slice) of every tensor - what the PS moves per id."""
        return [(t.numel() // max(t.shape[0], 1)) * t.element_size() if t.dim() > 0 else t.element_size() for t in self._tensors]
