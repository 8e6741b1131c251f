"""Guards used while a module is being traced / exported.

Reference: ``torchrec/pt2/checks.py`` (``set/get_use_torchdynamo_compiling_path`` :18-23, ``is_torchdynamo_compiling`` / ``is_non_strict_exporting`` :26-56,
``is_pt2_compiling`` :59, ``pt2_checks_tensor_slice`` :63, ``pt2_checks_all_is_size`` :77, ``pt2_check_size_nonzero`` :86, ``pt2_guard_size_oblivious`` :95).
``torch.compile`` is not this framework's hot path (CUDA graphs are); these checks exist so the export / IR tooling (``ir/``, ``fx/``) can run.
"""
from __future__ import annotations

from typing import List

import torch

USE_TORCHDYNAMO_COMPILING_PATH: bool = False


def set_use_torchdynamo_compiling_path(val: bool) -> None:
    global USE_TORCHDYNAMO_COMPILING_PATH
    USE_TORCHDYNAMO_COMPILING_PATH = val


def get_use_torchdynamo_compiling_path() -> bool:
    return USE_TORCHDYNAMO_COMPILING_PATH


def is_torchdynamo_compiling() -> bool:
    if USE_TORCHDYNAMO_COMPILING_PATH:
        return True
    try:
        return bool(torch.compiler.is_compiling())
    except Exception:
        return False


def is_non_strict_exporting() -> bool:
    try:
        return bool(torch.compiler.is_exporting()) and not torch.compiler.is_dynamo_compiling()
    except Exception:
        return False


def is_pt2_compiling() -> bool:
    return is_torchdynamo_compiling() or is_non_strict_exporting()


def pt2_checks_tensor_slice(tensor: torch.Tensor, start_offset: int, end_offset: int, dim: int = 0) -> None:
    """Tell the symbolic-shape engine that ``[start, end)`` is a valid slice of ``tensor`` along ``dim``."""
    if torch.jit.is_scripting() or not is_pt2_compiling():
        return
    torch._check_is_size(start_offset)
    torch._check_is_size(end_offset)
    torch._check_is_size(end_offset - start_offset)
    torch._check(start_offset <= tensor.size(dim))
    torch._check(end_offset <= tensor.size(dim))
    torch._check(end_offset >= start_offset)


def pt2_checks_all_is_size(x: List[int]) -> List[int]:
    if torch.jit.is_scripting() or not is_pt2_compiling():
        return x
    for i in x:
        torch._check_is_size(i)
    return x


def pt2_check_size_nonzero(x: torch.Tensor) -> torch.Tensor:
    if torch.jit.is_scripting() or not is_pt2_compiling():
        return x
    for i in range(x.dim()):
        torch._check(x.size(i) > 0)
    return x


def pt2_guard_size_oblivious(x: bool) -> bool:
    if torch.jit.is_scripting() or not is_pt2_compiling():
        return x
    try:
        from torch.fx.experimental.symbolic_shapes import guard_size_oblivious

        return guard_size_oblivious(x)
    except Exception:
        return x
