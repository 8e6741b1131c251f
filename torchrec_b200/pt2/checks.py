"""Hints for the symbolic-shape engine while a module is exported or traced.

Counterpart of ``torchrec/pt2/checks.py``. ``torch.compile`` is not a hot path of this framework (CUDA graphs are), so the only consumers are the export /
IR tooling (``ir/``, ``fx/``): every helper is the identity in eager mode and under TorchScript, and turns into ``torch._check*`` facts when a compile
or a non-strict export is in flight. The "are we compiling" probe can be forced on for tests with ``set_use_torchdynamo_compiling_path(True)``.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, TypeVar

import torch

T = TypeVar("T")
_forced_compiling = False


def set_use_torchdynamo_compiling_path(val: bool) -> None:
    global _forced_compiling
    _forced_compiling = bool(val)


def get_use_torchdynamo_compiling_path() -> bool:
    return _forced_compiling


def _probe(name: str) -> bool:
    fn = getattr(torch.compiler, name, None)
    try:
        return bool(fn()) if fn is not None else False
    except Exception:
        return False


def is_torchdynamo_compiling() -> bool:
    return _forced_compiling or _probe("is_compiling")


def is_non_strict_exporting() -> bool:
    return _probe("is_exporting") and not _probe("is_dynamo_compiling")


def is_pt2_compiling() -> bool:
    return is_torchdynamo_compiling() or is_non_strict_exporting()


def _hints_active() -> bool:
    return (not torch.jit.is_scripting()) and is_pt2_compiling()


def _all_sizes(values: Iterable[int]) -> None:
    for v in values:
        torch._check_is_size(v)


def pt2_checks_tensor_slice(tensor: torch.Tensor, start_offset: int, end_offset: int, dim: int = 0) -> None:
    """``tensor.narrow(dim, start, end - start)`` is in bounds: 0 <= start <= end <= size."""
    if _hints_active():
        _all_sizes((start_offset, end_offset, end_offset - start_offset))
        extent = tensor.size(dim)
        for fact in (start_offset <= extent, end_offset <= extent, end_offset >= start_offset):
            torch._check(fact)


def pt2_checks_all_is_size(x: List[int]) -> List[int]:
    if _hints_active():
        _all_sizes(x)
    return x


def pt2_check_size_nonzero(x: torch.Tensor) -> torch.Tensor:
    if _hints_active():
        for extent in x.shape:
            torch._check(extent > 0)
    return x


def pt2_guard_size_oblivious(x: bool) -> bool:
    if not _hints_active():
        return x
    try:
        from torch.fx.experimental.symbolic_shapes import guard_size_oblivious
    except ImportError:
        return x
    return guard_size_oblivious(x)
