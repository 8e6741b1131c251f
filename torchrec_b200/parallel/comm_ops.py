"""Autograd-aware asynchronous collectives for embedding redistribution.

Same op surface as the reference (torchrec/distributed/comm_ops.py:460-1340): ``alltoall_pooled``,
``variable_batch_alltoall_pooled``, ``alltoall_sequence``, ``alltoallv``, ``reduce_scatter_pooled``,
``reduce_scatter_base_pooled``, ``all_gather_base_pooled``, ``reduce_scatter_v_pooled``. Each returns
an ``Awaitable`` whose ``wait()`` yields a tensor wired into autograd; the backward launches the
mirror collective. Optional wire codecs (fp16/bf16/fp8/int8) wrap both directions.

This is the *portable* transport (NCCL on CUDA, Gloo on CPU) used for multi-node jobs, CPU tests
and as the internal baseline. On a single NVLink domain the sharded modules use the fused P2P
kernels in ``torchrec_b200.parallel.p2p`` instead, which need none of the pack/unpack steps here.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass, field
from typing import Any, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch.autograd import Function
from torch.autograd.profiler import record_function

from .types import Awaitable, NoWait, QuantizedCommCodecs

# Gradients of pooled embeddings are divided by the world size in backward so that sparse and
# (DDP-averaged) dense gradients see the same effective loss scaling (reference comm_ops.py:49-60).
GRADIENT_DIVISION: bool = True
USE_SYNC_COLLECTIVES: bool = False


def set_gradient_division(val: bool) -> None:
    global GRADIENT_DIVISION
    GRADIENT_DIVISION = val


def get_gradient_division() -> bool:
    return GRADIENT_DIVISION


def set_use_sync_collectives(val: bool) -> None:
    global USE_SYNC_COLLECTIVES
    USE_SYNC_COLLECTIVES = val


def get_use_sync_collectives() -> bool:
    return USE_SYNC_COLLECTIVES


def _pg(group: Optional[dist.ProcessGroup]) -> dist.ProcessGroup:
    return group if group is not None else dist.distributed_c10d._get_default_group()


class _Handle:
    """Mutable box shared by the launch / wait halves of one collective."""

    def __init__(self) -> None:
        self.work: Optional[dist.Work] = None
        self.buf: Optional[torch.Tensor] = None
        self.extra: Any = None


def _wait(h: _Handle) -> None:
    if h.work is not None:
        h.work.wait()
        h.work = None


class Request(Awaitable[torch.Tensor]):
    """Awaitable over a launched collective; ``wait()`` runs the completion half under autograd."""

    def __init__(self, finish) -> None:
        super().__init__()
        self._finish = finish

    def _wait_impl(self) -> torch.Tensor:
        out = self._finish()
        if out.is_cuda:
            out.record_stream(torch.cuda.current_stream(out.device))
        return out


# ---- pooled all-to-all --------------------------------------------------------------------------
@dataclass
class All2AllPooledInfo:
    batch_size_per_rank: List[int]
    dim_sum_per_rank: List[int]
    dim_sum_per_rank_tensor: Optional[torch.Tensor] = None
    cumsum_dim_sum_per_rank_tensor: Optional[torch.Tensor] = None
    codecs: Optional[QuantizedCommCodecs] = None
    comm: Optional[Any] = None  # All2AllSingle: allocation + transport of the forward exchange


def _enc(codecs: Optional[QuantizedCommCodecs], t: torch.Tensor, fwd: bool) -> torch.Tensor:
    if codecs is None:
        return t
    c = codecs.forward if fwd else codecs.backward
    return c.encode(t)


def _dec(codecs: Optional[QuantizedCommCodecs], t: torch.Tensor, fwd: bool, dtype: torch.dtype) -> torch.Tensor:
    if codecs is None:
        return t
    c = codecs.forward if fwd else codecs.backward
    out = c.decode(t)
    return out if out.dtype == dtype else out.to(dtype)


class _A2APooledReq(Function):
    @staticmethod
    def forward(ctx, pg, h: _Handle, info: All2AllPooledInfo, x: torch.Tensor) -> torch.Tensor:
        my_rank = dist.get_rank(pg)
        W = dist.get_world_size(pg)
        B_local = info.batch_size_per_rank[my_rank]
        D_local = info.dim_sum_per_rank[my_rank]
        ctx.pg, ctx.h, ctx.info, ctx.W = pg, h, info, W
        ctx.in_shape, ctx.in_dtype = x.shape, x.dtype
        send = _enc(info.codecs, x.contiguous().view(-1), True)
        scale = send.numel() // max(x.numel(), 1) if x.numel() else 1
        in_splits = [b * D_local for b in info.batch_size_per_rank]
        out_splits = [B_local * d for d in info.dim_sum_per_rank]
        if info.codecs is not None:
            in_splits = [info.codecs.forward.calc_quantized_size(s) for s in in_splits]
            out_splits = [info.codecs.forward.calc_quantized_size(s) for s in out_splits]
        comm = getattr(info, "comm", None)
        with record_function("## alltoall_pooled fwd ##"):
            if comm is not None:  # caller-provided allocation + transport (All2AllSingle)
                recv = comm.allocate(sum(out_splits), send.dtype, x.device)
                h.work = comm.all_to_all_single(recv, send, out_splits, in_splits, async_op=True)
            else:
                recv = torch.empty(sum(out_splits), dtype=send.dtype, device=x.device)
                h.work = dist.all_to_all_single(recv, send, out_splits, in_splits, group=pg, async_op=True)
        h.buf = recv
        h.extra = (out_splits, B_local, x.dtype)
        return x.new_zeros(1)

    @staticmethod
    def backward(ctx, _dummy_grad):
        h, info = ctx.h, ctx.info
        _wait(h)
        g = _dec(info.codecs, h.buf, False, ctx.in_dtype).view(ctx.in_shape)
        if GRADIENT_DIVISION:
            g = g / ctx.W
        h.buf = None
        return None, None, None, g


class _A2APooledWait(Function):
    @staticmethod
    def forward(ctx, pg, h: _Handle, info: All2AllPooledInfo, dummy: torch.Tensor) -> torch.Tensor:
        _wait(h)
        out_splits, B_local, dtype = h.extra
        ctx.pg, ctx.h, ctx.info = pg, h, info
        ctx.B_local = B_local
        parts = h.buf.split(out_splits)
        outs = [_dec(info.codecs, p, True, dtype).view(B_local, d) for p, d in zip(parts, info.dim_sum_per_rank)]
        h.buf = None
        return torch.cat(outs, dim=1)

    @staticmethod
    def backward(ctx, grad: torch.Tensor):
        pg, h, info = ctx.pg, ctx.h, ctx.info
        my_rank = dist.get_rank(pg)
        D_local = info.dim_sum_per_rank[my_rank]
        B_local = ctx.B_local
        pieces = [p.contiguous().view(-1) for p in grad.split(info.dim_sum_per_rank, dim=1)]
        send = _enc(info.codecs, torch.cat(pieces), False)
        in_splits = [B_local * d for d in info.dim_sum_per_rank]
        out_splits = [b * D_local for b in info.batch_size_per_rank]
        if info.codecs is not None:
            in_splits = [info.codecs.backward.calc_quantized_size(s) for s in in_splits]
            out_splits = [info.codecs.backward.calc_quantized_size(s) for s in out_splits]
        recv = torch.empty(sum(out_splits), dtype=send.dtype, device=grad.device)
        with record_function("## alltoall_pooled bwd ##"):
            h.work = dist.all_to_all_single(recv, send, out_splits, in_splits, group=pg, async_op=True)
        h.buf = recv
        return None, None, None, grad.new_zeros(1)


def alltoall_pooled(
    a2a_pooled_embs_tensor: torch.Tensor,
    batch_size_per_rank: List[int],
    dim_sum_per_rank: List[int],
    dim_sum_per_rank_tensor: Optional[torch.Tensor] = None,
    cumsum_dim_sum_per_rank_tensor: Optional[torch.Tensor] = None,
    group: Optional[dist.ProcessGroup] = None,
    codecs: Optional[QuantizedCommCodecs] = None,
    comm: Optional[Any] = None,
    all_to_all_single_comm: Optional[Any] = None,
) -> Awaitable[torch.Tensor]:
    """Pooled embeddings ``[sum_r B_r, D_local]`` (rows grouped by destination rank) ->
    ``[B_local, sum_r D_r]`` (columns grouped by source rank). ``comm``: an ``All2AllSingle`` (comm_ops_sync.py) that supplies
    the receive-buffer allocation and the transport of the forward exchange."""
    pg = _pg(group)
    if dist.get_world_size(pg) <= 1:
        return NoWait(a2a_pooled_embs_tensor)
    if USE_SYNC_COLLECTIVES and codecs is None:
        from .comm_ops_sync import all2all_pooled_sync

        return NoWait(all2all_pooled_sync(pg, batch_size_per_rank, dim_sum_per_rank, a2a_pooled_embs_tensor, None, GRADIENT_DIVISION))
    info = All2AllPooledInfo(batch_size_per_rank, dim_sum_per_rank, dim_sum_per_rank_tensor, cumsum_dim_sum_per_rank_tensor, codecs)
    info.comm = comm if comm is not None else all_to_all_single_comm  # ``all_to_all_single_comm``: the reference's name of the argument
    h = _Handle()
    dummy = _A2APooledReq.apply(pg, h, info, a2a_pooled_embs_tensor)
    return Request(lambda: _A2APooledWait.apply(pg, h, info, dummy))


# ---- variable batch pooled all-to-all -----------------------------------------------------------------
@dataclass
class VariableBatchAll2AllPooledInfo:
    batch_size_per_rank_per_feature: List[List[int]]
    batch_size_per_feature_pre_a2a: List[int]
    emb_dim_per_rank_per_feature: List[List[int]]
    codecs: Optional[QuantizedCommCodecs] = None
    input_splits: Optional[List[int]] = None
    output_splits: Optional[List[int]] = None


class _VBA2APooled(Function):
    @staticmethod
    def forward(ctx, pg, info: VariableBatchAll2AllPooledInfo, x: torch.Tensor) -> torch.Tensor:
        my_rank = dist.get_rank(pg)
        W = dist.get_world_size(pg)
        # send: for dest r, features local to me with batch b[r][f] -> sum_f b[r][f] * dim[me][f]
        in_splits = [sum(b * d for b, d in zip(info.batch_size_per_rank_per_feature[r], info.emb_dim_per_rank_per_feature[my_rank])) for r in range(W)]
        # recv from r: its features with my batch sizes
        bs_pre = info.batch_size_per_feature_pre_a2a
        out_splits, idx = [], 0
        for r in range(W):
            dims = info.emb_dim_per_rank_per_feature[r]
            out_splits.append(sum(b * d for b, d in zip(bs_pre[idx : idx + len(dims)], dims)))
            idx += len(dims)
        ctx.pg, ctx.W = pg, W
        ctx.in_splits, ctx.out_splits = in_splits, out_splits
        recv = torch.empty(sum(out_splits), dtype=x.dtype, device=x.device)
        dist.all_to_all_single(recv, x.contiguous(), out_splits, in_splits, group=pg)
        return recv

    @staticmethod
    def backward(ctx, grad):
        recv = torch.empty(sum(ctx.in_splits), dtype=grad.dtype, device=grad.device)
        dist.all_to_all_single(recv, grad.contiguous(), ctx.in_splits, ctx.out_splits, group=ctx.pg)
        if GRADIENT_DIVISION:
            recv = recv / ctx.W
        return None, None, recv


def variable_batch_alltoall_pooled(
    a2a_pooled_embs_tensor: torch.Tensor,
    batch_size_per_rank_per_feature: List[List[int]],
    batch_size_per_feature_pre_a2a: List[int],
    emb_dim_per_rank_per_feature: List[List[int]],
    group: Optional[dist.ProcessGroup] = None,
    codecs: Optional[QuantizedCommCodecs] = None,
    all_to_all_single_comm: Optional[Any] = None,
) -> Awaitable[torch.Tensor]:
    """1-D flattened variable-batch (VBE) pooled all-to-all (reference comm_ops.py:668-745)."""
    pg = _pg(group)
    if dist.get_world_size(pg) <= 1:
        return NoWait(a2a_pooled_embs_tensor)
    info = VariableBatchAll2AllPooledInfo(batch_size_per_rank_per_feature, batch_size_per_feature_pre_a2a, emb_dim_per_rank_per_feature, codecs)
    return NoWait(_VBA2APooled.apply(pg, info, a2a_pooled_embs_tensor))


# ---- sequence all-to-all ----------------------------------------------------------------------------------
@dataclass
class All2AllSequenceInfo:
    embedding_dim: int
    lengths_after_sparse_data_all2all: Optional[torch.Tensor]
    forward_recat_tensor: Optional[torch.Tensor]
    backward_recat_tensor: Optional[torch.Tensor]
    input_splits: List[int]
    output_splits: List[int]
    variable_batch_size: bool = False
    codecs: Optional[QuantizedCommCodecs] = None
    permuted_lengths_after_sparse_data_all2all: Optional[torch.Tensor] = None
    # variable batch size per rank: bags per (unit, source rank) segment of ``lengths_after_sparse_data_all2all`` (None: every segment
    # holds lengths.numel() / n_segments bags)
    bags_per_segment: Optional[torch.Tensor] = None


class _A2ASeqReq(Function):
    @staticmethod
    def forward(ctx, pg, h: _Handle, info: All2AllSequenceInfo, x: torch.Tensor) -> torch.Tensor:
        from ..ops import jagged as J

        D = info.embedding_dim
        W = dist.get_world_size(pg)
        ctx.pg, ctx.h, ctx.info, ctx.W = pg, h, info, W
        ctx.in_dtype = x.dtype
        if info.forward_recat_tensor is not None and info.lengths_after_sparse_data_all2all is not None and info.forward_recat_tensor.numel() > 0:
            # rows arrive unit-major; regroup them source-rank-major before sending them back (a rank without units has nothing to regroup)
            lengths = info.lengths_after_sparse_data_all2all
            nseg = info.forward_recat_tensor.numel()
            if info.bags_per_segment is not None:
                csum = torch.cat([lengths.new_zeros(1, dtype=torch.int64), torch.cumsum(lengths.to(torch.int64), 0)])
                ends = torch.cumsum(info.bags_per_segment.to(torch.int64), 0)
                seg = (csum[ends] - csum[ends - info.bags_per_segment.to(torch.int64)]).to(lengths.dtype)
            else:
                seg = lengths.view(nseg, -1).sum(1)
            _, x, _ = J.permute_1D_sparse_data(info.forward_recat_tensor, seg, x, None, x.shape[0])
            ctx.seg = seg
        else:
            ctx.seg = None
        send = _enc(info.codecs, x.contiguous().view(-1), True)
        in_splits = [s * D for s in info.input_splits]
        out_splits = [s * D for s in info.output_splits]
        if info.codecs is not None:
            in_splits = [info.codecs.forward.calc_quantized_size(s) for s in in_splits]
            out_splits = [info.codecs.forward.calc_quantized_size(s) for s in out_splits]
        recv = torch.empty(sum(out_splits), dtype=send.dtype, device=x.device)
        with record_function("## alltoall_seq_embedding fwd ##"):
            h.work = dist.all_to_all_single(recv, send, out_splits, in_splits, group=pg, async_op=True)
        h.buf = recv
        h.extra = x.dtype
        return x.new_zeros(1)

    @staticmethod
    def backward(ctx, _g):
        from ..ops import jagged as J

        h, info = ctx.h, ctx.info
        _wait(h)
        g = _dec(info.codecs, h.buf, False, ctx.in_dtype).view(-1, info.embedding_dim)
        h.buf = None
        if ctx.seg is not None and info.backward_recat_tensor is not None:
            seg_perm = ctx.seg[info.forward_recat_tensor.long()]
            _, g, _ = J.permute_1D_sparse_data(info.backward_recat_tensor, seg_perm, g, None, g.shape[0])
        if GRADIENT_DIVISION:  # same convention as the pooled all-to-all: every rank's loss is a local mean, the update is the global mean
            g = g / dist.get_world_size(ctx.pg)
        return None, None, None, g


class _A2ASeqWait(Function):
    @staticmethod
    def forward(ctx, pg, h: _Handle, info: All2AllSequenceInfo, dummy: torch.Tensor) -> torch.Tensor:
        _wait(h)
        ctx.pg, ctx.h, ctx.info = pg, h, info
        out = _dec(info.codecs, h.buf, True, h.extra).view(-1, info.embedding_dim)
        h.buf = None
        return out

    @staticmethod
    def backward(ctx, grad):
        pg, h, info = ctx.pg, ctx.h, ctx.info
        D = info.embedding_dim
        send = _enc(info.codecs, grad.contiguous().view(-1), False)
        in_splits = [s * D for s in info.output_splits]
        out_splits = [s * D for s in info.input_splits]
        if info.codecs is not None:
            in_splits = [info.codecs.backward.calc_quantized_size(s) for s in in_splits]
            out_splits = [info.codecs.backward.calc_quantized_size(s) for s in out_splits]
        recv = torch.empty(sum(out_splits), dtype=send.dtype, device=grad.device)
        with record_function("## alltoall_seq_embedding bwd ##"):
            h.work = dist.all_to_all_single(recv, send, out_splits, in_splits, group=pg, async_op=True)
        h.buf = recv
        return None, None, None, grad.new_zeros(1)


def alltoall_sequence(
    a2a_sequence_embs_tensor: torch.Tensor,
    forward_recat_tensor: Optional[torch.Tensor],
    backward_recat_tensor: Optional[torch.Tensor],
    lengths_after_sparse_data_all2all: Optional[torch.Tensor],
    input_splits: List[int],
    output_splits: List[int],
    variable_batch_size: bool = False,
    group: Optional[dist.ProcessGroup] = None,
    codecs: Optional[QuantizedCommCodecs] = None,
    batch_size_per_rank: Optional[List[int]] = None,
) -> Awaitable[torch.Tensor]:
    """Send unpooled embedding rows ``[sum L, D]`` back to the ranks that own the samples. ``batch_size_per_rank``: the ranks fed
    different batch sizes (the received lengths are [unit][source rank][that rank's batch])."""
    pg = _pg(group)
    if dist.get_world_size(pg) <= 1:
        return NoWait(a2a_sequence_embs_tensor)
    info = All2AllSequenceInfo(a2a_sequence_embs_tensor.shape[1], lengths_after_sparse_data_all2all, forward_recat_tensor,
                               backward_recat_tensor, input_splits, output_splits, variable_batch_size, codecs)
    if batch_size_per_rank is not None and len(set(batch_size_per_rank)) > 1 and forward_recat_tensor is not None and forward_recat_tensor.numel() > 0:
        W = len(batch_size_per_rank)
        n_units = forward_recat_tensor.numel() // W
        info.bags_per_segment = torch.tensor(list(batch_size_per_rank) * n_units, dtype=torch.int64, device=a2a_sequence_embs_tensor.device)
        info.variable_batch_size = True
    h = _Handle()
    dummy = _A2ASeqReq.apply(pg, h, info, a2a_sequence_embs_tensor)
    return Request(lambda: _A2ASeqWait.apply(pg, h, info, dummy))


# ---- all-to-all-v -------------------------------------------------------------------------------------------
class _A2AV(Function):
    @staticmethod
    def forward(ctx, pg, out_split: List[int], per_rank_split_lengths: Optional[List[int]], *inputs: torch.Tensor):
        my_rank = dist.get_rank(pg)
        W = dist.get_world_size(pg)
        B_global = inputs[0].shape[0]
        dims = [t.shape[1] for t in inputs]
        D_local = sum(dims)
        if per_rank_split_lengths is None:
            B_local = B_global // W
            per_rank_split_lengths = [B_local] * W
        B_local = per_rank_split_lengths[my_rank]
        x = torch.cat(inputs, dim=1).contiguous()
        in_splits = [b * D_local for b in per_rank_split_lengths]
        out_splits = [B_local * d for d in out_split]
        recv = torch.empty(sum(out_splits), dtype=x.dtype, device=x.device)
        dist.all_to_all_single(recv, x.view(-1), out_splits, in_splits, group=pg)
        ctx.pg, ctx.W, ctx.dims = pg, W, dims
        ctx.in_splits, ctx.out_splits, ctx.B_global, ctx.out_split, ctx.B_local = in_splits, out_splits, B_global, out_split, B_local
        return tuple(p.view(B_local, d) for p, d in zip(recv.split(out_splits), out_split))

    @staticmethod
    def backward(ctx, *grads):
        send = torch.cat([g.contiguous().view(-1) for g in grads])
        recv = torch.empty(sum(ctx.in_splits), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send, ctx.in_splits, ctx.out_splits, group=ctx.pg)
        g = recv.view(ctx.B_global, -1)  # the generic all-to-all-v carries gradients unscaled (only the embedding dists divide by W)
        return (None, None, None) + tuple(g.split(ctx.dims, dim=1))


def alltoallv(inputs: List[torch.Tensor], out_split: Optional[List[int]] = None, per_rank_split_lengths: Optional[List[int]] = None,
              group: Optional[dist.ProcessGroup] = None, codecs: Optional[QuantizedCommCodecs] = None) -> Awaitable[List[torch.Tensor]]:
    """All-to-all of a list of ``[B_global, D_i]`` tensors; returns one ``[B_local, out_split[r]]`` per rank."""
    pg = _pg(group)
    if dist.get_world_size(pg) <= 1:
        return NoWait(list(inputs))
    assert out_split is not None, "alltoallv needs the embedding dim sum of every rank (out_split)"
    return NoWait(list(_A2AV.apply(pg, out_split, per_rank_split_lengths, *inputs)))


# ---- reduce-scatter / all-gather family ---------------------------------------------------------------------------
def _reduce_scatter_sum(pg, x: torch.Tensor, in_rows: List[int]) -> torch.Tensor:
    """Sum-reduce-scatter along dim 0 with (possibly uneven) row splits. Implemented as an
    all-to-all of the slices followed by a local sum: same bytes on the wire, works on every
    backend (Gloo has no reduce_scatter) and matches the NVSwitch-friendly P2P schedule."""
    W = dist.get_world_size(pg)
    my_rank = dist.get_rank(pg)
    cols = x.shape[1:] if x.dim() > 1 else ()
    n_my = in_rows[my_rank]
    width = int(torch.tensor(cols).prod()) if len(cols) else 1
    in_splits = [r * width for r in in_rows]
    out_splits = [n_my * width] * W
    recv = torch.empty(W * n_my * width, dtype=x.dtype, device=x.device)
    dist.all_to_all_single(recv, x.contiguous().view(-1), out_splits, in_splits, group=pg)
    return recv.view(W, n_my, *cols).sum(0)


def _all_gather_rows(pg, x: torch.Tensor, rows: List[int]) -> torch.Tensor:
    W = dist.get_world_size(pg)
    if all(r == rows[0] for r in rows):
        out = torch.empty((W * rows[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=pg)
        return out
    outs = [torch.empty((r,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device) for r in rows]
    dist.all_gather(outs, x.contiguous(), group=pg)
    return torch.cat(outs, 0)


class _ReduceScatterV(Function):
    @staticmethod
    def forward(ctx, pg, in_rows: List[int], codecs, x: torch.Tensor) -> torch.Tensor:
        ctx.pg, ctx.in_rows, ctx.codecs = pg, in_rows, codecs
        ctx.W = dist.get_world_size(pg)
        dtype = x.dtype
        xs = x
        if codecs is not None:
            e = codecs.forward.encode(x)
            if e.shape == x.shape:  # only shape-preserving (cast) codecs can be summed on the wire
                xs = e
        out = _reduce_scatter_sum(pg, xs, in_rows)
        return out.to(dtype) if out.dtype != dtype else out

    @staticmethod
    def backward(ctx, grad):
        g = grad
        if ctx.codecs is not None:
            ge = ctx.codecs.backward.encode(grad)
            if ge.shape == grad.shape:
                g = ge
        full = _all_gather_rows(ctx.pg, g, ctx.in_rows).to(grad.dtype)
        if GRADIENT_DIVISION:
            full = full / ctx.W
        return None, None, None, full


def reduce_scatter_v_pooled(input: torch.Tensor, input_splits: List[int], group: Optional[dist.ProcessGroup] = None,
                            codecs: Optional[QuantizedCommCodecs] = None) -> Awaitable[torch.Tensor]:
    """Sum-reduce ``[sum_r B_r, D]`` over ranks and scatter rows with uneven ``input_splits``."""
    pg = _pg(group)
    if dist.get_world_size(pg) <= 1:
        return NoWait(input)
    if USE_SYNC_COLLECTIVES and codecs is None:
        from .comm_ops_sync import reduce_scatter_v_sync

        return NoWait(reduce_scatter_v_sync(pg, input, list(input_splits), GRADIENT_DIVISION))
    return NoWait(_ReduceScatterV.apply(pg, list(input_splits), codecs, input))


def reduce_scatter_base_pooled(input: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                               codecs: Optional[QuantizedCommCodecs] = None) -> Awaitable[torch.Tensor]:
    """``[W * B, D]`` -> ``[B, D]`` (sum over ranks); backward is an all-gather."""
    pg = _pg(group)
    W = dist.get_world_size(pg)
    if W <= 1:
        return NoWait(input)
    if USE_SYNC_COLLECTIVES and codecs is None:
        from .comm_ops_sync import reduce_scatter_base_sync

        return NoWait(reduce_scatter_base_sync(pg, input, GRADIENT_DIVISION))
    return NoWait(_ReduceScatterV.apply(pg, [input.shape[0] // W] * W, codecs, input))


def reduce_scatter_pooled(inputs: List[torch.Tensor], group: Optional[dist.ProcessGroup] = None,
                          codecs: Optional[QuantizedCommCodecs] = None) -> Awaitable[torch.Tensor]:
    """List form: ``inputs[r]`` is the slab destined to rank r."""
    pg = _pg(group)
    if dist.get_world_size(pg) <= 1:
        return NoWait(inputs[0])
    rows = [t.shape[0] for t in inputs]
    return NoWait(_ReduceScatterV.apply(pg, rows, codecs, torch.cat(inputs, 0)))


def reduce_scatter_v_per_feature_pooled(input: torch.Tensor, batch_size_per_rank_per_feature: List[List[int]], embedding_dims: List[int],
                                        group: Optional[dist.ProcessGroup] = None, codecs: Optional[QuantizedCommCodecs] = None) -> Awaitable[torch.Tensor]:
    """VBE reduce-scatter over a 1-D buffer laid out [rank][feature][batch, dim]."""
    pg = _pg(group)
    if dist.get_world_size(pg) <= 1:
        return NoWait(input)
    sizes = [sum(b * d for b, d in zip(bs, embedding_dims)) for bs in batch_size_per_rank_per_feature]
    return NoWait(_ReduceScatterV.apply(pg, sizes, codecs, input.view(-1)))


class _AllGatherBase(Function):
    @staticmethod
    def forward(ctx, pg, codecs, x: torch.Tensor) -> torch.Tensor:
        ctx.pg = pg
        ctx.W = dist.get_world_size(pg)
        ctx.rows = x.shape[0]
        return _all_gather_rows(pg, x, [x.shape[0]] * ctx.W)

    @staticmethod
    def backward(ctx, grad):
        g = _reduce_scatter_sum(ctx.pg, grad, [ctx.rows] * ctx.W)
        if GRADIENT_DIVISION:
            g = g / ctx.W
        return None, None, g


def all_gather_base_pooled(input: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                           codecs: Optional[QuantizedCommCodecs] = None) -> Awaitable[torch.Tensor]:
    """``[B, D]`` -> ``[W * B, D]``; backward is a reduce-scatter."""
    pg = _pg(group)
    if dist.get_world_size(pg) <= 1:
        return NoWait(input)
    if USE_SYNC_COLLECTIVES and codecs is None:
        from .comm_ops_sync import all_gather_base_sync

        return NoWait(all_gather_base_sync(pg, input, GRADIENT_DIVISION))
    return NoWait(_AllGatherBase.apply(pg, codecs, input))


# ---- reference surface: argument records of the collectives, sync-mode switch, names of the sync transports ---------------------------------------------
@contextlib.contextmanager
def torchrec_use_sync_collectives():
    """Inside the block the functional (traceable, blocking) collectives of ``comm_ops_sync`` are used instead of the request / wait pair."""
    original = get_use_sync_collectives()
    set_use_sync_collectives(True)
    try:
        yield
    finally:
        set_use_sync_collectives(original)


def pg_name(pg: dist.ProcessGroup) -> str:
    """The registered name of a process group (what the functional collectives take)."""
    return dist._functional_collectives._resolve_group_name(pg, "")


@dataclass
class All2AllVInfo:
    """Arguments of ``alltoallv``: column counts per rank, global / local batch sizes, per-table batch sizes and dims, split sizes."""

    dims_sum_per_rank: List[int]
    B_global: int
    B_local: int
    B_local_list: List[int]
    D_local_list: List[int]
    input_split_sizes: List[int] = field(default_factory=list)
    output_split_sizes: List[int] = field(default_factory=list)
    codecs: Optional[QuantizedCommCodecs] = None


@dataclass
class ReduceScatterInfo:
    input_sizes: List[torch.Size]
    codecs: Optional[QuantizedCommCodecs] = None


@dataclass
class ReduceScatterBaseInfo:
    input_sizes: torch.Size
    codecs: Optional[QuantizedCommCodecs] = None


@dataclass
class AllGatherBaseInfo:
    input_size: torch.Size
    codecs: Optional[QuantizedCommCodecs] = None


@dataclass
class ReduceScatterVInfo:
    input_sizes: List[List[int]]
    input_splits: List[int]
    equal_splits: bool
    total_input_size: List[int]
    codecs: Optional[QuantizedCommCodecs] = None


@dataclass
class All2AllDenseInfo:
    output_splits: List[int]
    batch_size: int
    input_shape: List[int]
    input_splits: List[int]


def __getattr__(name: str):
    """The blocking functional collectives and their transports live in ``comm_ops_sync``; they are importable from here under the
    reference's names."""
    if name in ("all2all_pooled_sync", "all2all_sequence_sync", "reduce_scatter_base_sync", "all_gather_base_sync", "reduce_scatter_v_sync",
                "variable_batch_all2all_pooled_sync", "reduce_scatter_tensor", "all_gather_into_tensor", "Comm", "All2AllSingle", "DefaultAll2AllSingle"):
        from . import comm_ops_sync as _m

        return getattr(_m, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
