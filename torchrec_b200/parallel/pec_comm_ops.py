"""All-to-all of sequence embeddings for prioritized embedding communication (PEC).

Reference: ``torchrec/distributed/pec_comm_ops.py`` - ``PECAll2AllSeqInfo`` :24, ``_grad_dist`` :63, ``PECAll2AllSeqWait`` :106. PEC splits a sequence
lookup into an early (non-overlapped, needed first) and a late part; both go through this op so their backward collectives can be issued
independently and the early one never waits for the late one.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .types import Awaitable


@dataclass
class PECAll2AllSeqInfo:
    """Split sizes (in ROWS) of one PEC sequence all-to-all and the permutation that restores the requester's row order."""

    input_splits: List[int]
    output_splits: List[int]
    embedding_dim: int
    permute: Optional[torch.Tensor] = None
    pg: Optional[dist.ProcessGroup] = None


def _a2a_rows(x: torch.Tensor, in_splits: List[int], out_splits: List[int], pg: Optional[dist.ProcessGroup]) -> Tuple[torch.Tensor, Optional[dist.Work]]:
    out = torch.empty(sum(out_splits), x.shape[1], dtype=x.dtype, device=x.device)
    if pg is None or dist.get_world_size(pg) == 1:
        out.copy_(x)
        return out, None
    work = dist.all_to_all_single(out, x.contiguous(), output_split_sizes=out_splits, input_split_sizes=in_splits, group=pg, async_op=True)
    return out, work


def _grad_dist(grad: torch.Tensor, info: PECAll2AllSeqInfo) -> torch.Tensor:
    """Backward of the forward all-to-all: the same exchange with the splits swapped (un-permute first)."""
    if info.permute is not None:
        inv = torch.empty_like(info.permute)
        inv[info.permute] = torch.arange(info.permute.numel(), device=info.permute.device)
        grad = grad[inv]
    out, work = _a2a_rows(grad, info.output_splits, info.input_splits, info.pg)
    if work is not None:
        work.wait()
    return out


class PECAll2AllSeqWait(torch.autograd.Function):
    """Differentiable wait: forward waits the async all-to-all (and applies ``permute``), backward runs the reverse exchange."""

    @staticmethod
    def forward(ctx, anchor: torch.Tensor, out: torch.Tensor, work, info: PECAll2AllSeqInfo) -> torch.Tensor:  # type: ignore[override]
        if work is not None:
            work.wait()
        ctx.info = info
        return out[info.permute] if info.permute is not None else out

    @staticmethod
    def backward(ctx, grad: torch.Tensor):  # type: ignore[override]
        return _grad_dist(grad.contiguous(), ctx.info), None, None, None


class PECAll2AllSeqAwaitable(Awaitable[torch.Tensor]):
    def __init__(self, local_embs: torch.Tensor, info: PECAll2AllSeqInfo) -> None:
        super().__init__()
        self._embs, self._info = local_embs, info
        self._out, self._work = _a2a_rows(local_embs.detach(), info.input_splits, info.output_splits, info.pg)

    def _wait_impl(self) -> torch.Tensor:
        return PECAll2AllSeqWait.apply(self._embs, self._out, self._work, self._info)


def pec_all2all_seq(local_embs: torch.Tensor, info: PECAll2AllSeqInfo) -> Awaitable[torch.Tensor]:
    """Launch the exchange now, differentiate through ``.wait()`` later."""
    return PECAll2AllSeqAwaitable(local_embs, info)
