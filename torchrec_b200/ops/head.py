"""Fused CTR head: Linear(K -> 1) on bf16 activations and BCE-with-logits (mean) with its gradient in one pass
(csrc/head.cu). Falls back to plain PyTorch on CPU."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib


class RowDotFn(torch.autograd.Function):
    """``logits[b] = <x[b], w> + bias`` for bf16 ``x [B, K]`` and fp32 ``w [1, K]``. If ``x`` is the ReLU output of a fused
    layer (``_trb_relu_out``), that layer's ReLU mask is applied inside this backward (``_trb_masked`` on the gradient)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
        B, K = x.shape
        out = torch.empty(B, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        w = weight.detach().reshape(-1).contiguous()
        code = L.trb_rowdot_fwd(_lib.ptr(x), ctypes.c_int64(x.stride(0)), _lib.ptr(w), _lib.ptr(bias.detach() if bias is not None else None), _lib.ptr(out), B, K,
                                _lib.stream_ptr(x.device))
        _lib.check(code, "trb_rowdot_fwd")
        ctx.mask = bool(getattr(x, "_trb_relu_out", False))
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w)
        return out.unsqueeze(1)

    @staticmethod
    def backward(ctx, gy: torch.Tensor):
        x, w = ctx.saved_tensors
        B, K = x.shape
        dz = gy.reshape(-1).float().contiguous()
        dx = torch.empty(B, K, dtype=torch.bfloat16, device=x.device) if ctx.needs_input_grad[0] else None
        dw = torch.empty(K, dtype=torch.float32, device=x.device)
        db = torch.empty(1, dtype=torch.float32, device=x.device) if ctx.has_bias else None
        L = _lib.lib()
        code = L.trb_rowdot_bwd(_lib.ptr(x), ctypes.c_int64(x.stride(0)), _lib.ptr(dz), _lib.ptr(w), _lib.ptr(dx), ctypes.c_int64(K), _lib.ptr(dw), _lib.ptr(db), B, K,
                                int(ctx.mask), ctypes.c_float(1.0), _lib.stream_ptr(x.device))
        _lib.check(code, "trb_rowdot_bwd")
        if dx is not None and ctx.mask:
            dx._trb_masked = True
        return dx, dw.view(1, K), db


def rowdot_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and weight.shape[0] == 1 and weight.dtype == torch.float32 and x.shape[1] % 8 == 0
            and x.shape[1] <= 2048 and x.stride(1) == 1 and x.stride(0) % 8 == 0)


class BCEWithLogitsMeanFn(torch.autograd.Function):
    """mean BCE-with-logits; the gradient w.r.t. the logits is produced by the same kernel as the loss."""

    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        z = logits.reshape(-1).float().contiguous()
        B = z.numel()
        kind = {torch.float32: 0, torch.int64: 1, torch.int32: 2}.get(labels.dtype)
        if kind is None:
            labels, kind = labels.float(), 0
        labels = labels.reshape(-1).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z)
        L = _lib.lib()
        code = L.trb_bce_fwd_bwd(_lib.ptr(z), _lib.ptr(labels), kind, _lib.ptr(loss), _lib.ptr(dz), B, _lib.stream_ptr(z.device))
        _lib.check(code, "trb_bce_fwd_bwd")
        ctx.save_for_backward(dz)
        ctx.shape = logits.shape
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        (dz,) = ctx.saved_tensors
        return (dz * g).view(ctx.shape), None


def bce_with_logits_mean(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    if logits.is_cuda and _lib.available():
        return BCEWithLogitsMeanFn.apply(logits, labels)
    return torch.nn.functional.binary_cross_entropy_with_logits(logits.float(), labels.float())
