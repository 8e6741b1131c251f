"""PT2 helpers (reference torchrec/pt2/utils.py, pt2/checks.py). torch.compile is NOT a hot path of this framework (the
kernels are hand written); these helpers only keep user code that guards on dynamo working."""
from typing import List

import torch

from ..sparse.jagged_tensor import KeyedJaggedTensor


def is_torchdynamo_compiling() -> bool:
    try:
        return torch.compiler.is_compiling()
    except AttributeError:
        return False


def is_non_strict_exporting() -> bool:
    try:
        return torch.compiler.is_exporting()
    except AttributeError:
        return False


def is_pt2_compiling() -> bool:
    return is_torchdynamo_compiling() or is_non_strict_exporting()


def pt2_checks_tensor_slice(tensor: torch.Tensor, start_offset: int, end_offset: int, dim: int = 0) -> None:
    if not is_pt2_compiling():
        return
    torch._check_is_size(start_offset)
    torch._check_is_size(end_offset)
    torch._check(start_offset <= end_offset)
    torch._check(end_offset <= tensor.size(dim))


def pt2_checks_all_is_size(x: List[int]) -> List[int]:
    if is_pt2_compiling():
        for i in x:
            torch._check_is_size(i)
    return x


def pt2_check_size_nonzero(x: torch.Tensor) -> torch.Tensor:
    if is_pt2_compiling():
        for i in range(x.dim()):
            torch._check(x.size(i) > 0)
    return x


def kjt_for_pt2_tracing(kjt: KeyedJaggedTensor, convert_to_vb: bool = False, mark_length: bool = False) -> KeyedJaggedTensor:
    """A KJT whose host-side caches are dropped and whose tensors are marked dynamic along the value dimension."""
    out = KeyedJaggedTensor(keys=kjt.keys(), values=kjt.values(), lengths=kjt.lengths().long(), weights=kjt.weights_or_none(), stride=kjt.stride() if not convert_to_vb else None,
                            stride_per_key_per_rank=[[kjt.stride()]] * len(kjt.keys()) if convert_to_vb else None)
    try:
        torch._dynamo.mark_dynamic(out.values(), 0)
        if out.weights_or_none() is not None:
            torch._dynamo.mark_dynamic(out.weights(), 0)
        if mark_length:
            torch._dynamo.mark_dynamic(out.lengths(), 0)
    except Exception:
        pass
    return out
