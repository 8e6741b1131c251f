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


def default_pipeline_input_transformer(inp):
    """Input hook of ``TrainPipelinePT2``: the KJT fields of a batch are re-made for tracing (host caches dropped, value dimension dynamic)."""
    for attr_name in ("id_list_features", "id_score_list_features", "sparse_features"):
        attr = getattr(inp, attr_name, None)
        if isinstance(attr, KeyedJaggedTensor):
            setattr(inp, attr_name, kjt_for_pt2_tracing(attr))
    return inp


def pt2_compile_callable(f):
    """Decorator for metric ``update`` / ``compute`` methods: when the owning object sets ``enable_pt2_compile = True`` the method runs
    through ``torch.compile`` (compiled once per object and cached on it), otherwise it is called as is. Opt-in only: nothing in this
    framework's hot path depends on a tracing compiler."""
    import functools

    attr = f"_{f.__name__}_pt2_compiled"

    @functools.wraps(f)
    def inner(ref, *args, **kwargs):
        if getattr(ref, "enable_pt2_compile", False):
            compiled = ref.__dict__.get(attr) if hasattr(ref, "__dict__") else None
            if compiled is None:
                compiled = torch.compile(f)
                object.__setattr__(ref, attr, compiled)
            return compiled(ref, *args, **kwargs)
        return f(ref, *args, **kwargs)

    return inner


class AtomicCounter:
    """Python stand-in of the ``fbgemm::AtomicCounter`` script class some reference models hold (per-table step counters): same
    methods, plain integer. Nothing to register with the fake-class registry - it is an ordinary Python object under tracing too."""

    def __init__(self, counter_: int = 0) -> None:
        self.counter_ = int(counter_)

    def increment(self) -> int:
        self.counter_ += 1
        return self.counter_

    def decrement(self) -> int:
        self.counter_ -= 1
        return self.counter_

    def reset(self) -> None:
        self.counter_ = 0

    def get(self) -> int:
        return self.counter_

    def set(self, val: int) -> None:
        self.counter_ = int(val)


class TensorQueue:
    """Python stand-in of ``fbgemm::TensorQueue``: FIFO of tensors, ``pop`` / ``top`` of an empty queue return the initial tensor."""

    def __init__(self, init_tensor: torch.Tensor, queue=None) -> None:
        self.init_tensor = init_tensor
        self.queue = list(queue or [])

    def push(self, x: torch.Tensor) -> None:
        self.queue.append(x)

    def pop(self) -> torch.Tensor:
        return self.queue.pop(0) if self.queue else self.init_tensor

    def top(self) -> torch.Tensor:
        return self.queue[0] if self.queue else self.init_tensor

    def size(self) -> int:
        return len(self.queue)


def register_fake_classes() -> None:
    """The reference registers fake (meta) versions of the two FBGEMM script classes above so they trace under PT2. Here they are Python
    classes already; kept as a no-op so call sites keep working."""
    return None


def deregister_fake_classes() -> None:
    return None
