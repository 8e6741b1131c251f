"""Synchronous, traceable variants of the embedding collectives + the pluggable all-to-all hook.

``comm_ops.set_use_sync_collectives(True)`` switches ``alltoall_pooled`` / ``alltoall_sequence`` / ``reduce_scatter_base_pooled`` /
``all_gather_base_pooled`` / ``reduce_scatter_v_pooled`` / ``variable_batch_alltoall_pooled`` to the implementations here: plain functions over
``torch.distributed._functional_collectives`` (results are ready when the call returns; no Awaitable state machine, no custom autograd.Function holding
a process group), which is what graph capture / export tooling needs. The three ops without a functional counterpart are registered as custom ops
with fake (shape-only) implementations and autograd formulas:

    torchrec_b200::reduce_scatter_tensor(input, reduceOp, group_size, group_name, gradient_division) -> Tensor        backward = all_gather
    torchrec_b200::all_gather_into_tensor(shard, gather_dim, group_size, group_name, gradient_division) -> Tensor     backward = reduce_scatter
    torchrec_b200::_split_1d_cat_2d(tensor, dim0, dim1_splits) -> Tensor    flat per-rank blocks [sum_r B x D_r] -> [B, sum_r D_r]

``All2AllSingle`` is the hook through which a caller supplies its own buffer allocation + all-to-all for the pooled forward exchange (e.g. NCCL
user-buffer registration, or this repository's NVLink peer-memory transport); ``DefaultAll2AllSingle`` is ``torch.empty`` + ``dist.all_to_all_single``.

Capability parity: reference comm_ops.py ``Comm`` / ``All2AllSingle`` / ``DefaultAll2AllSingle`` :89-126, 383-457; sync functions and custom ops :2678-2884.
"""
from __future__ import annotations

import abc
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.distributed._functional_collectives as fc

from .types import QuantizedCommCodecs


# ---- pluggable all-to-all ---------------------------------------------------------------------------------------------------------------
class Comm(abc.ABC):
    """A communication provider bound to one process group."""

    def __init__(self, pg: Optional[dist.ProcessGroup] = None) -> None:
        self.pg = pg


class All2AllSingle(Comm):
    """Allocation + transport of the pooled forward all-to-all (``alltoall_pooled(..., comm=...)``)."""

    @abc.abstractmethod
    def allocate(self, numel: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor: ...

    @abc.abstractmethod
    def all_to_all_single(self, output: torch.Tensor, input: torch.Tensor, output_split_sizes: Sequence[int], input_split_sizes: Sequence[int],
                          async_op: bool = True) -> Optional[dist.Work]: ...


class DefaultAll2AllSingle(All2AllSingle):
    def allocate(self, numel: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        return torch.empty(numel, dtype=dtype, device=device)

    def all_to_all_single(self, output, input, output_split_sizes, input_split_sizes, async_op: bool = True):
        return dist.all_to_all_single(output, input, list(output_split_sizes), list(input_split_sizes), group=self.pg, async_op=async_op)


# ---- custom ops ---------------------------------------------------------------------------------------------------------------------------
def _group_of(name: str) -> dist.ProcessGroup:
    from torch.distributed.distributed_c10d import _resolve_process_group

    return _resolve_process_group(name)


@torch.library.custom_op("torchrec_b200::reduce_scatter_tensor", mutates_args=())
def reduce_scatter_tensor(input: torch.Tensor, reduceOp: str, group_size: int, group_name: str, gradient_division: bool) -> torch.Tensor:
    out = fc.reduce_scatter_tensor(input, reduceOp, 0, _group_of(group_name))
    return fc.wait_tensor(out) if hasattr(fc, "wait_tensor") else out


@reduce_scatter_tensor.register_fake
def _(input, reduceOp, group_size, group_name, gradient_division):
    return input.new_empty((input.shape[0] // group_size,) + tuple(input.shape[1:]))


def _rs_setup(ctx, inputs, output):
    _, _, ctx.group_size, ctx.group_name, ctx.gradient_division = inputs


def _rs_backward(ctx, grad):
    g = all_gather_into_tensor(grad.contiguous(), 0, ctx.group_size, ctx.group_name, False)
    if ctx.gradient_division:
        g = g / ctx.group_size
    return g, None, None, None, None


reduce_scatter_tensor.register_autograd(_rs_backward, setup_context=_rs_setup)


@torch.library.custom_op("torchrec_b200::all_gather_into_tensor", mutates_args=())
def all_gather_into_tensor(shard: torch.Tensor, gather_dim: int, group_size: int, group_name: str, gradient_division: bool) -> torch.Tensor:
    out = fc.all_gather_tensor(shard, gather_dim, _group_of(group_name))
    return fc.wait_tensor(out) if hasattr(fc, "wait_tensor") else out


@all_gather_into_tensor.register_fake
def _(shard, gather_dim, group_size, group_name, gradient_division):
    shape = list(shard.shape)
    shape[gather_dim] *= group_size
    return shard.new_empty(shape)


def _ag_setup(ctx, inputs, output):
    _, ctx.gather_dim, ctx.group_size, ctx.group_name, ctx.gradient_division = inputs


def _ag_backward(ctx, grad):
    assert ctx.gather_dim == 0, "all_gather_into_tensor backward is defined for gather_dim 0"
    g = reduce_scatter_tensor(grad.contiguous(), "sum", ctx.group_size, ctx.group_name, False)
    if ctx.gradient_division:
        g = g / ctx.group_size
    return g, None, None, None, None


all_gather_into_tensor.register_autograd(_ag_backward, setup_context=_ag_setup)


@torch.library.custom_op("torchrec_b200::_split_1d_cat_2d", mutates_args=())
def _split_1d_cat_2d(tensor: torch.Tensor, dim0: int, dim1_splits: List[int]) -> torch.Tensor:
    parts = torch.split(tensor, [dim0 * d for d in dim1_splits])
    return torch.cat([p.view(dim0, d) for p, d in zip(parts, dim1_splits)], dim=1)


@_split_1d_cat_2d.register_fake
def _(tensor, dim0, dim1_splits):
    return tensor.new_empty((dim0, sum(dim1_splits)))


def _sc_setup(ctx, inputs, output):
    _, ctx.dim0, ctx.dim1_splits = inputs


def _sc_backward(ctx, grad):
    return torch.cat([g.contiguous().view(-1) for g in torch.split(grad, ctx.dim1_splits, dim=1)]), None, None


_split_1d_cat_2d.register_autograd(_sc_backward, setup_context=_sc_setup)


# ---- sync collectives ------------------------------------------------------------------------------------------------------------------------
def _enc(codecs: Optional[QuantizedCommCodecs], t: torch.Tensor, fwd: bool) -> torch.Tensor:
    return t if codecs is None else (codecs.forward if fwd else codecs.backward).encode(t)


def _dec(codecs: Optional[QuantizedCommCodecs], t: torch.Tensor, fwd: bool, dtype: torch.dtype) -> torch.Tensor:
    if codecs is None:
        return t
    out = (codecs.forward if fwd else codecs.backward).decode(t)
    return out if out.dtype == dtype else out.to(dtype)


def _a2a_flat(pg: dist.ProcessGroup, send: torch.Tensor, out_splits: List[int], in_splits: List[int]) -> torch.Tensor:
    """Differentiable all-to-all of a flat tensor (functional collective: the backward is the transposed exchange)."""
    return fc.all_to_all_single_autograd(send, out_splits, in_splits, pg)


def all2all_pooled_sync(pg: dist.ProcessGroup, batch_size_per_rank: List[int], dim_sum_per_rank: List[int], x: torch.Tensor,
                        codecs: Optional[QuantizedCommCodecs] = None, gradient_division: bool = True) -> torch.Tensor:
    """``[sum_r B_r, D_local]`` -> ``[B_local, sum_r D_r]``; gradients flow through the functional all-to-all."""
    me, W = dist.get_rank(pg), dist.get_world_size(pg)
    B_local, D_local = batch_size_per_rank[me], dim_sum_per_rank[me]
    in_splits = [b * D_local for b in batch_size_per_rank]
    out_splits = [B_local * d for d in dim_sum_per_rank]
    if codecs is not None:
        # wire codecs are not differentiable: quantised exchanges keep the request / wait functions of comm_ops (straight-through there)
        raise NotImplementedError("sync collectives carry fp32 / bf16 tensors; use the default (async) ops with quantized comm codecs")
    recv = _a2a_flat(pg, x.contiguous().view(-1), out_splits, in_splits)
    out = _split_1d_cat_2d(recv, B_local, list(dim_sum_per_rank))
    if gradient_division and x.requires_grad:
        out = _GradScale.apply(out, 1.0 / W)
    return out


def all2all_sequence_sync(pg: dist.ProcessGroup, x: torch.Tensor, input_splits: List[int], output_splits: List[int], gradient_division: bool = True) -> torch.Tensor:
    """Unpooled rows ``[sum L, D]`` back to the sample owners (rows already grouped by destination rank)."""
    D = x.shape[1]
    recv = _a2a_flat(pg, x.contiguous().view(-1), [s * D for s in output_splits], [s * D for s in input_splits]).view(-1, D)
    if gradient_division and x.requires_grad:
        recv = _GradScale.apply(recv, 1.0 / dist.get_world_size(pg))
    return recv


def reduce_scatter_base_sync(pg: dist.ProcessGroup, x: torch.Tensor, gradient_division: bool = True) -> torch.Tensor:
    """``[W * B, D]`` -> summed ``[B, D]`` (row-wise pooled output dist)."""
    return reduce_scatter_tensor(x.contiguous(), "sum", dist.get_world_size(pg), pg.group_name, gradient_division)


def all_gather_base_sync(pg: dist.ProcessGroup, x: torch.Tensor, gradient_division: bool = True) -> torch.Tensor:
    return all_gather_into_tensor(x.contiguous(), 0, dist.get_world_size(pg), pg.group_name, gradient_division)


def reduce_scatter_v_sync(pg: dist.ProcessGroup, x: torch.Tensor, input_splits: List[int], gradient_division: bool = True) -> torch.Tensor:
    """Uneven row blocks: rank r receives the sum of every rank's block r (``input_splits[r]`` rows). Expressed as a differentiable
    all-to-all of the blocks followed by a local sum over the source ranks."""
    me, W = dist.get_rank(pg), dist.get_world_size(pg)
    D = x.shape[1]
    recv = _a2a_flat(pg, x.contiguous().view(-1), [input_splits[me] * D] * W, [s * D for s in input_splits])
    out = recv.view(W, input_splits[me], D).sum(0)
    if gradient_division and x.requires_grad:
        out = _GradScale.apply(out, 1.0 / W)
    return out


def variable_batch_all2all_pooled_sync(pg: dist.ProcessGroup, x: torch.Tensor, batch_size_per_rank_per_feature: List[List[int]],
                                        batch_size_per_feature_pre_a2a: List[int], emb_dim_per_rank_per_feature: List[List[int]],
                                        gradient_division: bool = True) -> torch.Tensor:
    """Variable batch per feature: 1-D buffers, splits = sum over the rank's features of (batch x dim)."""
    me, W = dist.get_rank(pg), dist.get_world_size(pg)
    in_splits = [sum(b * d for b, d in zip(batch_size_per_rank_per_feature[r], emb_dim_per_rank_per_feature[me])) for r in range(W)]
    out_splits, off = [], 0
    for r in range(W):
        n = len(emb_dim_per_rank_per_feature[r])
        out_splits.append(sum(b * d for b, d in zip(batch_size_per_feature_pre_a2a[off : off + n], emb_dim_per_rank_per_feature[r])))
        off += n
    recv = _a2a_flat(pg, x.contiguous().view(-1), out_splits, in_splits)
    if gradient_division and x.requires_grad:
        recv = _GradScale.apply(recv, 1.0 / W)
    return recv


class _GradScale(torch.autograd.Function):
    """Identity forward, gradient x ``scale`` backward (the 1 / W division of the embedding dists)."""

    @staticmethod
    def forward(ctx, x, scale: float):
        ctx.scale = scale
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None
