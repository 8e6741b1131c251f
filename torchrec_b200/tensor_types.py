"""Packed sub-byte tensors (reference torchrec/tensor_types.py:61-190): ``UInt4Tensor`` / ``UInt2Tensor`` are ``torch.Tensor`` wrapper
subclasses over a uint8 storage holding 2 / 4 elements per byte - the logical view of INT4 / INT2 quantized embedding rows (without the
fused scale / bias tail), so state-dict tooling sees shapes in elements, ``isinstance(t, torch.Tensor)`` holds and the tensor survives
``detach`` / ``clone`` / ``copy_`` / ``to`` / row and column slicing / ``view(torch.uint8)``. Element k of a byte sits in bits
[k * bits, (k + 1) * bits) (low bits first); ``pack`` / ``unpack`` convert from / to one-element-per-entry uint8 tensors."""
from typing import Any, List, Tuple

import torch


def down_size(N: int, size: torch.Size) -> Tuple[int, ...]:
    """Shape of the byte storage of a logical shape with ``N`` elements per byte."""
    assert size[-1] % N == 0, f"{size} last dim not divisible by {N}"
    return (*size[:-1], size[-1] // N)


def up_size(N: int, size: torch.Size) -> Tuple[int, ...]:
    """Logical shape of a byte storage with ``N`` elements per byte."""
    return (*size[:-1], size[-1] * N)


def fill_defaults(args, n: int, defaults_tail):
    """``__torch_dispatch__`` does not pass defaulted arguments: pad ``args`` to ``n`` entries with the tail of ``defaults_tail``
    (``fill_defaults([1, 2, 3], 5, [3, 4, 5]) == [1, 2, 3, 4, 5]``)."""
    if n - len(defaults_tail) > len(args):
        raise RuntimeError("not enough defaults to fill arguments")
    r = list(args)
    for i in range(len(args), n):
        r.append(defaults_tail[i - n + len(defaults_tail)])
    return r


def find_arg_of_type(it, t):
    for x in it:
        if isinstance(x, t):
            return x
    return None


class UIntXTensor(torch.Tensor):
    """``N`` elements per byte of the wrapped uint8 tensor ``elem``; the last dimension is the packed one."""

    __torch_function__ = torch._C._disabled_torch_function_impl
    BITS = 8

    @staticmethod
    def __new__(cls, N: int, elem: torch.Tensor):
        assert elem.dtype is torch.uint8, "packed storage must be uint8"
        return torch.Tensor._make_wrapper_subclass(cls, up_size(N, elem.shape), dtype=torch.uint8, device=elem.device)

    def __init__(self, N: int, elem: torch.Tensor) -> None:
        self.N = N
        self.elem = elem

    # ---- element access -------------------------------------------------------------------------------------------------------------
    @property
    def per_byte(self) -> int:
        return self.N

    def unpack(self) -> torch.Tensor:
        """uint8 tensor of the logical shape, one element per entry."""
        bits = 8 // self.N
        mask = (1 << bits) - 1
        parts = [(self.elem >> (bits * k)) & mask for k in range(self.N)]
        return torch.stack(parts, dim=-1).reshape(*self.elem.shape[:-1], -1)

    def tolist(self) -> List:
        return self.unpack().tolist()

    def __repr__(self) -> str:
        return f"{type(self).__name__}(shape={tuple(self.shape)}, device={self.elem.device}, elem={self.elem})"

    def _wrap(self, elem: torch.Tensor) -> "UIntXTensor":
        return type(self)(elem) if type(self) is not UIntXTensor else UIntXTensor(self.N, elem)

    @classmethod
    def __torch_dispatch__(cls, func, types, args, kwargs=None):  # noqa: C901
        kwargs = kwargs or {}
        aten = torch.ops.aten
        self = find_arg_of_type(args, UIntXTensor)
        if func is aten.detach.default:
            return self._wrap(self.elem.detach())
        if func is aten.clone.default:
            return self._wrap(self.elem.clone())
        if func is aten.copy_.default:
            dst, src = args[0], args[1]
            dst.elem.copy_(src.elem if isinstance(src, UIntXTensor) else src)
            return dst
        if func is aten.view.dtype:
            (dtype,) = args[1:]
            if dtype == torch.uint8:  # the bytes
                return self.elem
            raise NotImplementedError(f"view of a packed tensor as {dtype}")
        if func is aten._to_copy.default:
            dtype = kwargs.get("dtype")
            if dtype is not None and dtype != torch.uint8:
                raise NotImplementedError(f"packed tensors cannot be cast to {dtype}")
            kw = {k: v for k, v in kwargs.items() if k in ("device", "non_blocking", "memory_format", "pin_memory", "layout") and v is not None}
            return self._wrap(aten._to_copy.default(self.elem, **kw))
        if func is aten.slice.Tensor:
            t, dim, start, end, step = fill_defaults(args, 5, [0, None, None, 1])
            nd = t.dim()
            dim = dim % nd
            if dim == nd - 1:  # the packed dimension: whole bytes only
                size = t.shape[-1]
                start = 0 if start is None else max(min(start + size if start < 0 else start, size), 0)
                end = size if end is None else max(min(end + size if end < 0 else end, size), 0)
                if step != 1 or start % t.N or end % t.N:
                    raise NotImplementedError(f"column slice [{start}:{end}:{step}] of a packed tensor must cover whole bytes ({t.N} elements each)")
                return t._wrap(aten.slice.Tensor(t.elem, dim, start // t.N, end // t.N, 1))
            return t._wrap(aten.slice.Tensor(t.elem, dim, start, end, step))
        if func is aten.select.int:
            t, dim, index = args
            if dim % t.dim() != t.dim() - 1:
                return t._wrap(aten.select.int(t.elem, dim, index))
            raise NotImplementedError("select along the packed dimension")
        if func is aten.view.default:
            t, size = args
            size = list(size)
            if len(size) == 1 and size[0] == -1:
                size = [t.numel()]
            return t._wrap(t.elem.reshape(down_size(t.N, torch.Size(size))))
        if func is aten.alias.default:
            return self._wrap(self.elem)
        raise NotImplementedError(f"{func} on {cls.__name__}")

    # ---- construction from unpacked values ----------------------------------------------------------------------------------------------------
    @classmethod
    def pack(cls, values: torch.Tensor):
        per = 8 // cls.BITS
        bits = cls.BITS
        assert values.shape[-1] % per == 0, f"last dim {values.shape[-1]} not divisible by {per}"
        v = values.to(torch.uint8).reshape(*values.shape[:-1], -1, per)
        out = torch.zeros(v.shape[:-1], dtype=torch.uint8, device=values.device)
        for k in range(per):
            out |= (v[..., k] & ((1 << bits) - 1)) << (bits * k)
        return cls(out)


class UInt4Tensor(UIntXTensor):
    BITS = 4
    N: int = 2

    @staticmethod
    def __new__(cls, elem: torch.Tensor):
        return UIntXTensor.__new__(cls, 2, elem)

    def __init__(self, elem: torch.Tensor) -> None:
        super().__init__(2, elem)


class UInt2Tensor(UIntXTensor):
    BITS = 2
    N: int = 4

    @staticmethod
    def __new__(cls, elem: torch.Tensor):
        return UIntXTensor.__new__(cls, 4, elem)

    def __init__(self, elem: torch.Tensor) -> None:
        super().__init__(4, elem)
