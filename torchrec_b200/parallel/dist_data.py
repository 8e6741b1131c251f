"""Redistribution modules: KJT all-to-all (input dist) and embedding dists (output dist).

Module surface of the reference (torchrec/distributed/dist_data.py): ``KJTAllToAll`` (two-stage
awaitable: splits exchange -> tensors exchange -> recat), ``PooledEmbeddingsAllToAll``,
``PooledEmbeddingsReduceScatter``, ``SequenceEmbeddingsAllToAll``, variable-batch variants,
``PooledEmbeddingsAllGather``, ``TensorAllToAll``/``JaggedTensorAllToAll`` and the single-process
inference movers (``KJTOneToAll``, ``EmbeddingsAllToOne``...). Portable NCCL/Gloo transport; the
single-NVLink-domain fast path is in ``torchrec_b200.parallel.p2p``.
"""
from __future__ import annotations

import itertools
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn
from torch.autograd.profiler import record_function

from ..ops import jagged as J
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .comm_ops import (
    all_gather_base_pooled,
    alltoall_pooled,
    alltoall_sequence,
    reduce_scatter_base_pooled,
    reduce_scatter_v_per_feature_pooled,
    reduce_scatter_v_pooled,
    variable_batch_alltoall_pooled,
)
from .types import Awaitable, NoWait, QuantizedCommCodecs

# debug switches (reference dist_data.py:105-112)
TORCHREC_OVERFLOW_DEBUG = os.environ.get("TORCHREC_OVERFLOW_DEBUG", "0") == "1"
TORCHREC_VALIDATE_COLLECTIVES = os.environ.get("TORCHREC_VALIDATE_COLLECTIVES", "0") == "1"


def _fnv1a(s: str) -> int:
    h = 0x811C9DC5
    for ch in s.encode():
        h = ((h ^ ch) * 0x01000193) & 0x7FFFFFFF
    return h


def _get_recat(local_split: int, num_splits: int, stagger: int = 1, device: Optional[torch.device] = None,
               batch_size_per_rank: Optional[List[int]] = None) -> Optional[torch.Tensor]:
    """Permutation that turns received [rank][key] segments into [key][rank] order. ``stagger`` > 1
    interleaves ranks node-wise (used by table-row-wise style dists). With uneven batch sizes the
    caller permutes segments (1-D), so the recat stays at segment granularity."""
    if local_split == 0:
        return None
    recat: List[int] = []
    feature_order: List[int] = [x + num_splits // stagger * y for x in range(num_splits // stagger) for y in range(stagger)]
    for i in range(local_split):
        for j in feature_order:
            recat.append(i + j * local_split)
    return torch.tensor(recat, device=device, dtype=torch.int32)


class SplitsAllToAllAwaitable(Awaitable[List[List[int]]]):
    """Exchange per-tensor split sizes (small int all-to-all, read back on the host)."""

    def __init__(self, input_tensors: List[torch.Tensor], pg: dist.ProcessGroup, tag: Optional[int] = None) -> None:
        super().__init__()
        self.num_workers = pg.size()
        self._tag = tag
        rows = list(input_tensors)
        if tag is not None:
            rows = rows + [torch.full((self.num_workers,), tag, dtype=rows[0].dtype, device=rows[0].device)]
        self._n = len(rows)
        with record_function("## all2all_data:kjt splits ##"):
            inp = torch.stack(rows, 1).flatten().contiguous()
            self._output_tensor = torch.empty_like(inp)
            self._work = dist.all_to_all_single(self._output_tensor, inp, group=pg, async_op=True)

    def _wait_impl(self) -> List[List[int]]:
        self._work.wait()
        rows = self._output_tensor.view(self.num_workers, -1).T.tolist()  # host sync (D2H)
        if self._tag is not None:
            tags = rows.pop()
            if any(t != self._tag for t in tags):
                raise RuntimeError(f"collective mismatch: this rank's input-dist tag {self._tag} differs from peers {tags}; "
                                   "ranks are executing different sharded modules / batches")
        return rows


class KJTAllToAllTensorsAwaitable(Awaitable[KeyedJaggedTensor]):
    """Second stage: one all-to-all per KJT tensor (lengths, values, [strides], [weights]), then
    the recat permute into key-major order."""

    def __init__(self, pg: dist.ProcessGroup, input: KeyedJaggedTensor, splits: List[int], input_splits: List[List[int]],
                 output_splits: List[List[int]], input_tensors: List[torch.Tensor], labels: List[str], keys: List[str],
                 device: torch.device, stagger: int, stride_per_rank: Optional[List[int]]) -> None:
        super().__init__()
        self._workers = pg.size()
        self._pg = pg
        self._device = device
        self._input = input
        self._splits = splits
        self._input_splits = input_splits
        self._output_splits = output_splits
        self._keys = keys
        self._stagger = stagger
        self._stride_per_rank = stride_per_rank
        self._recat = _get_recat(splits[pg.rank()], self._workers, stagger, device, stride_per_rank)
        self._output_tensors: List[torch.Tensor] = []
        self._awaitables: List[dist.Work] = []
        if self._workers == 1:
            return
        for in_t, in_s, out_s, label in zip(input_tensors, input_splits, output_splits, labels):
            out_t = torch.empty(sum(out_s), device=device, dtype=in_t.dtype)
            with record_function(f"## all2all_data:kjt {label} ##"):
                w = dist.all_to_all_single(out_t, in_t.contiguous(), output_split_sizes=out_s, input_split_sizes=in_s, group=pg, async_op=True)
            self._output_tensors.append(out_t)
            self._awaitables.append(w)

    def _wait_impl(self) -> KeyedJaggedTensor:
        if self._workers == 1:
            self._input.sync()
            return self._input
        for w in self._awaitables:
            w.wait()
        return KeyedJaggedTensor.dist_init(
            keys=self._keys, tensors=self._output_tensors, variable_stride_per_key=self._input.variable_stride_per_key(),
            num_workers=self._workers, recat=self._recat, stride_per_rank=self._stride_per_rank, stagger=self._stagger,
        )


class KJTAllToAllSplitsAwaitable(Awaitable[KJTAllToAllTensorsAwaitable]):
    """First stage: exchange split sizes; ``wait()`` launches the tensor all-to-alls."""

    def __init__(self, pg: dist.ProcessGroup, input: KeyedJaggedTensor, splits: List[int], labels: List[str],
                 tensor_splits: List[List[int]], input_tensors: List[torch.Tensor], keys: List[str], device: torch.device, stagger: int) -> None:
        super().__init__()
        self._workers = pg.size()
        self._pg = pg
        self._device = device
        self._input = input
        self._splits = splits
        self._labels = labels
        self._input_splits = tensor_splits
        self._input_tensors = input_tensors
        self._keys = keys
        self._stagger = stagger
        self._output_splits: List[List[int]] = tensor_splits
        self._stride_per_rank: Optional[List[int]] = None if input.variable_stride_per_key() else [input.stride()] * self._workers
        if self._workers == 1:
            self._splits_awaitable = None
            return
        input_tensors_splits = [torch.tensor(s, device=device, dtype=torch.int64) for s in tensor_splits]
        if not input.variable_stride_per_key():
            input_tensors_splits.append(torch.tensor([input.stride()] * self._workers, device=device, dtype=torch.int64))
        tag = _fnv1a("|".join(keys) + "|".join(labels)) if TORCHREC_VALIDATE_COLLECTIVES else None
        self._splits_awaitable = SplitsAllToAllAwaitable(input_tensors_splits, pg, tag)

    def _wait_impl(self) -> KJTAllToAllTensorsAwaitable:
        if self._workers > 1:
            output_list = self._splits_awaitable.wait()
            if not self._input.variable_stride_per_key():
                self._stride_per_rank = output_list.pop()
            self._output_splits = output_list
            if TORCHREC_OVERFLOW_DEBUG:
                for s in self._output_splits:
                    if sum(s) >= 2**31:
                        raise OverflowError(f"input-dist split sizes overflow int32: {s}")
        return KJTAllToAllTensorsAwaitable(
            pg=self._pg, input=self._input, splits=self._splits, input_splits=self._input_splits, output_splits=self._output_splits,
            input_tensors=self._input_tensors, labels=self._labels, keys=self._keys, device=self._device, stagger=self._stagger,
            stride_per_rank=self._stride_per_rank,
        )


class KJTAllToAll(nn.Module):
    """Redistribute a KJT so that rank r receives the keys ``[sum(splits[:r]), sum(splits[:r+1]))``
    of every rank's batch. ``forward`` returns ``Awaitable[Awaitable[KJT]]``."""

    def __init__(self, pg: dist.ProcessGroup, splits: List[int], stagger: int = 1) -> None:
        super().__init__()
        assert len(splits) == pg.size()
        self._pg = pg
        self._splits = splits
        self._splits_cumsum: List[int] = [0] + list(itertools.accumulate(splits))
        self._stagger = stagger

    def forward(self, input: KeyedJaggedTensor) -> Awaitable[KJTAllToAllTensorsAwaitable]:
        with torch.no_grad():
            assert len(input.keys()) == sum(self._splits)
            rank = dist.get_rank(self._pg)
            local_keys = input.keys()[self._splits_cumsum[rank] : self._splits_cumsum[rank + 1]]
            return KJTAllToAllSplitsAwaitable(
                pg=self._pg, input=input, splits=self._splits, labels=input.dist_labels(), tensor_splits=input.dist_splits(self._splits),
                input_tensors=input.dist_tensors(), keys=local_keys, device=input.device(), stagger=self._stagger,
            )


class KJTOneToAll(nn.Module):
    """Single-process inference: split a KJT by keys and copy piece r to device r (P2P copies)."""

    def __init__(self, splits: List[int], world_size: int, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._splits = splits
        self._world_size = world_size
        self._device_type = "meta" if device is not None and device.type == "meta" else ("cuda" if torch.cuda.is_available() else "cpu")
        assert self._world_size == len(splits)

    def forward(self, kjt: KeyedJaggedTensor) -> Awaitable[List[KeyedJaggedTensor]]:
        with torch.no_grad():
            pieces = kjt.split(self._splits)
            dist_kjts = [p if self._device_type != "cuda" else p.to(torch.device("cuda", r), non_blocking=True) for r, p in enumerate(pieces)]
            return NoWait(dist_kjts)


# ---- output dists -------------------------------------------------------------------------------------
class PooledEmbeddingsAwaitable(Awaitable[torch.Tensor]):
    def __init__(self, tensor_awaitable: Awaitable[torch.Tensor]) -> None:
        super().__init__()
        self._tensor_awaitable = tensor_awaitable

    def _wait_impl(self) -> torch.Tensor:
        ret = self._tensor_awaitable.wait()
        return ret

    @property
    def callbacks(self) -> List[Callable[[torch.Tensor], torch.Tensor]]:
        return self._callbacks


class PooledEmbeddingsAllToAll(nn.Module):
    """``[B_global, D_local]`` -> ``[B_local, sum_r D_r]`` (table-wise / column-wise output dist)."""

    def __init__(self, pg: dist.ProcessGroup, dim_sum_per_rank: List[int], device: Optional[torch.device] = None,
                 callbacks: Optional[List[Callable[[torch.Tensor], torch.Tensor]]] = None, codecs: Optional[QuantizedCommCodecs] = None) -> None:
        super().__init__()
        self._pg = pg
        self._callbacks: List[Callable[[torch.Tensor], torch.Tensor]] = list(callbacks) if callbacks is not None else []
        self._dim_sum_per_rank = dim_sum_per_rank
        self._codecs = codecs
        self.register_buffer("_dim_sum_per_rank_tensor", torch.tensor(dim_sum_per_rank, device=device, dtype=torch.int))
        cumsum = list(itertools.accumulate([0] + dim_sum_per_rank))
        self.register_buffer("_cumsum_dim_sum_per_rank_tensor", torch.tensor(cumsum, device=device, dtype=torch.int))

    def forward(self, local_embs: torch.Tensor, batch_size_per_rank: Optional[List[int]] = None) -> PooledEmbeddingsAwaitable:
        W = dist.get_world_size(self._pg)
        if batch_size_per_rank is None:
            B_global = local_embs.size(0)
            assert B_global % W == 0, f"num of ranks {W} doesn't divide global batch size {B_global}"
            batch_size_per_rank = [B_global // W] * W
        aw = alltoall_pooled(local_embs, batch_size_per_rank, self._dim_sum_per_rank, self._dim_sum_per_rank_tensor,
                             self._cumsum_dim_sum_per_rank_tensor, group=self._pg, codecs=self._codecs)
        out = PooledEmbeddingsAwaitable(aw)
        out.callbacks.extend(self._callbacks)
        return out

    @property
    def callbacks(self) -> List[Callable[[torch.Tensor], torch.Tensor]]:
        return self._callbacks


class VariableBatchPooledEmbeddingsAllToAll(nn.Module):
    def __init__(self, pg: dist.ProcessGroup, emb_dim_per_rank_per_feature: List[List[int]], device: Optional[torch.device] = None,
                 callbacks: Optional[List[Callable[[torch.Tensor], torch.Tensor]]] = None, codecs: Optional[QuantizedCommCodecs] = None) -> None:
        super().__init__()
        self._pg = pg
        self._emb_dim_per_rank_per_feature = emb_dim_per_rank_per_feature
        self._callbacks = list(callbacks) if callbacks is not None else []
        self._codecs = codecs

    def forward(self, local_embs: torch.Tensor, batch_size_per_rank_per_feature: List[List[int]], batch_size_per_feature_pre_a2a: List[int]) -> PooledEmbeddingsAwaitable:
        aw = variable_batch_alltoall_pooled(local_embs, batch_size_per_rank_per_feature, batch_size_per_feature_pre_a2a,
                                            self._emb_dim_per_rank_per_feature, group=self._pg, codecs=self._codecs)
        out = PooledEmbeddingsAwaitable(aw)
        out.callbacks.extend(self._callbacks)
        return out

    @property
    def callbacks(self):
        return self._callbacks


class PooledEmbeddingsReduceScatter(nn.Module):
    """``[sum_r B_r, D]`` -> ``[B_local, D]`` summed over ranks (row-wise output dist)."""

    def __init__(self, pg: dist.ProcessGroup, codecs: Optional[QuantizedCommCodecs] = None) -> None:
        super().__init__()
        self._pg = pg
        self._codecs = codecs

    def forward(self, local_embs: torch.Tensor, input_splits: Optional[List[int]] = None) -> PooledEmbeddingsAwaitable:
        if input_splits and len(set(input_splits)) > 1:
            aw = reduce_scatter_v_pooled(local_embs, input_splits, self._pg, codecs=self._codecs)
        else:
            aw = reduce_scatter_base_pooled(local_embs, self._pg, codecs=self._codecs)
        return PooledEmbeddingsAwaitable(aw)


class VariableBatchPooledEmbeddingsReduceScatter(nn.Module):
    def __init__(self, pg: dist.ProcessGroup, codecs: Optional[QuantizedCommCodecs] = None) -> None:
        super().__init__()
        self._pg = pg
        self._codecs = codecs

    def forward(self, local_embs: torch.Tensor, batch_size_per_rank_per_feature: List[List[int]], embedding_dims: List[int]) -> PooledEmbeddingsAwaitable:
        aw = reduce_scatter_v_per_feature_pooled(local_embs, batch_size_per_rank_per_feature, embedding_dims, self._pg, self._codecs)
        return PooledEmbeddingsAwaitable(aw)


class PooledEmbeddingsAllGather(nn.Module):
    def __init__(self, pg: dist.ProcessGroup, codecs: Optional[QuantizedCommCodecs] = None) -> None:
        super().__init__()
        self._pg = pg
        self._codecs = codecs

    def forward(self, local_emb: torch.Tensor) -> PooledEmbeddingsAwaitable:
        return PooledEmbeddingsAwaitable(all_gather_base_pooled(local_emb, self._pg, codecs=self._codecs))


class SequenceEmbeddingsAwaitable(Awaitable[torch.Tensor]):
    def __init__(self, tensor_awaitable: Awaitable[torch.Tensor], unbucketize_permute_tensor: Optional[torch.Tensor], embedding_dim: int) -> None:
        super().__init__()
        self._tensor_awaitable = tensor_awaitable
        self._unbucketize_permute_tensor = unbucketize_permute_tensor
        self._embedding_dim = embedding_dim
        if unbucketize_permute_tensor is not None and unbucketize_permute_tensor.dtype not in (torch.int32, torch.int64):
            raise ValueError("unbucketize_permute_tensor must be an integer tensor")

    def _wait_impl(self) -> torch.Tensor:
        ret = self._tensor_awaitable.wait()
        if self._unbucketize_permute_tensor is not None:
            ret = torch.index_select(ret.view(-1, self._embedding_dim), 0, self._unbucketize_permute_tensor.long())
        return ret


class SequenceEmbeddingsAllToAll(nn.Module):
    """Send unpooled embeddings back to the sample owners (inverse of the KJT all-to-all)."""

    def __init__(self, pg: dist.ProcessGroup, features_per_rank: List[int], device: Optional[torch.device] = None,
                 codecs: Optional[QuantizedCommCodecs] = None) -> None:
        super().__init__()
        self._pg = pg
        self._local_split = features_per_rank[pg.rank()]
        self._num_splits = pg.size()
        self._codecs = codecs
        fwd = []
        for j in range(self._num_splits):
            for i in range(self._local_split):
                fwd.append(j + i * self._num_splits)
        self.register_buffer("_forward_recat_tensor", torch.tensor(fwd, device=device, dtype=torch.int))
        self.register_buffer("_backward_recat_tensor", J.invert_permute(self._forward_recat_tensor) if fwd else self._forward_recat_tensor)

    def forward(self, local_embs: torch.Tensor, lengths: torch.Tensor, input_splits: List[int], output_splits: List[int],
                unbucketize_permute_tensor: Optional[torch.Tensor] = None, batch_size_per_rank: Optional[List[int]] = None,
                sparse_features_recat: Optional[torch.Tensor] = None) -> SequenceEmbeddingsAwaitable:
        variable_batch_size = batch_size_per_rank is not None and len(set(batch_size_per_rank)) > 1
        fwd = self._forward_recat_tensor if sparse_features_recat is None else J.invert_permute(sparse_features_recat)
        bwd = self._backward_recat_tensor if sparse_features_recat is None else sparse_features_recat
        aw = alltoall_sequence(local_embs, fwd, bwd, lengths, input_splits, output_splits, variable_batch_size, group=self._pg, codecs=self._codecs,
                               batch_size_per_rank=batch_size_per_rank)
        return SequenceEmbeddingsAwaitable(aw, unbucketize_permute_tensor, local_embs.shape[1])


# ---- generic tensor movers --------------------------------------------------------------------------------
class TensorAllToAllValuesAwaitable(Awaitable[torch.Tensor]):
    def __init__(self, pg: dist.ProcessGroup, input: torch.Tensor, input_splits: List[int], output_splits: List[int], device: torch.device) -> None:
        super().__init__()
        self._workers = pg.size()
        self._input = input
        self._dist_values = input
        self._work = None
        if self._workers > 1:
            shape = (sum(output_splits),) + tuple(input.shape[1:])
            self._dist_values = torch.empty(shape, device=device, dtype=input.dtype)
            w = 1
            for s in input.shape[1:]:
                w *= s
            self._work = dist.all_to_all_single(self._dist_values.view(-1), input.contiguous().view(-1), [o * w for o in output_splits],
                                                [i * w for i in input_splits], group=pg, async_op=True)

    def _wait_impl(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
        return self._dist_values


class TensorValuesAllToAll(nn.Module):
    def __init__(self, pg: dist.ProcessGroup) -> None:
        super().__init__()
        self._pg = pg

    def forward(self, input: torch.Tensor, input_splits, output_splits) -> TensorAllToAllValuesAwaitable:
        ins = input_splits.tolist() if isinstance(input_splits, torch.Tensor) else list(input_splits)
        outs = output_splits.tolist() if isinstance(output_splits, torch.Tensor) else list(output_splits)
        with torch.no_grad():
            return TensorAllToAllValuesAwaitable(self._pg, input, ins, outs, input.device)


class TensorAllToAllSplitsAwaitable(Awaitable[TensorAllToAllValuesAwaitable]):
    def __init__(self, pg: dist.ProcessGroup, input: torch.Tensor, splits: List[int], device: torch.device) -> None:
        super().__init__()
        self._pg = pg
        self._input = input
        self._splits = splits
        self._device = device
        self._aw = SplitsAllToAllAwaitable([torch.tensor(splits, device=device, dtype=torch.int64)], pg) if pg.size() > 1 else None

    def _wait_impl(self) -> TensorAllToAllValuesAwaitable:
        out_splits = self._aw.wait()[0] if self._aw is not None else self._splits
        return TensorAllToAllValuesAwaitable(self._pg, self._input, self._splits, out_splits, self._device)


class TensorAllToAll(nn.Module):
    """Redistribute rows of a tensor with splits known only to the sender."""

    def __init__(self, pg: dist.ProcessGroup) -> None:
        super().__init__()
        self._pg = pg

    def forward(self, input: torch.Tensor, splits: List[int]) -> TensorAllToAllSplitsAwaitable:
        with torch.no_grad():
            return TensorAllToAllSplitsAwaitable(self._pg, input, splits, input.device)


class JaggedTensorAllToAll(Awaitable[JaggedTensor]):
    """Redistribute a JaggedTensor given the number of *items* (bags) sent to each rank."""

    def __init__(self, jt: JaggedTensor, num_items_to_send: torch.Tensor, num_items_to_receive: torch.Tensor, pg: dist.ProcessGroup) -> None:
        super().__init__()
        self._workers = pg.size()
        self._dist_lengths = torch.empty(int(num_items_to_receive.sum()), device=jt.lengths().device, dtype=jt.lengths().dtype)
        send_items = num_items_to_send.tolist()
        recv_items = num_items_to_receive.tolist()
        dist.all_to_all_single(self._dist_lengths, jt.lengths().contiguous(), recv_items, send_items, group=pg)
        off = J.asynchronous_complete_cumsum(jt.lengths().long())
        bounds = torch.tensor([0] + list(itertools.accumulate(send_items)), device=off.device)
        send_vals = (off[bounds[1:]] - off[bounds[:-1]]).tolist()
        roff = J.asynchronous_complete_cumsum(self._dist_lengths.long())
        rbounds = torch.tensor([0] + list(itertools.accumulate(recv_items)), device=off.device)
        recv_vals = (roff[rbounds[1:]] - roff[rbounds[:-1]]).tolist()
        self._dist_values = torch.empty((sum(recv_vals),) + tuple(jt.values().shape[1:]), device=jt.values().device, dtype=jt.values().dtype)
        w = 1
        for s in jt.values().shape[1:]:
            w *= s
        dist.all_to_all_single(self._dist_values.view(-1), jt.values().contiguous().view(-1), [r * w for r in recv_vals], [s * w for s in send_vals], group=pg)
        self._dist_weights = None
        if jt.weights_or_none() is not None:
            self._dist_weights = torch.empty(sum(recv_vals), device=jt.values().device, dtype=jt.weights().dtype)
            dist.all_to_all_single(self._dist_weights, jt.weights().contiguous(), recv_vals, send_vals, group=pg)

    def _wait_impl(self) -> JaggedTensor:
        return JaggedTensor(values=self._dist_values, lengths=self._dist_lengths, weights=self._dist_weights)


# ---- single-process inference gathers (reference dist_data.py:387-420, 1555-1697) -----------------------------
class MergePooledEmbeddingsModule(nn.Module):
    """Concatenate per-device pooled embeddings onto one device (NVLink P2P copies)."""

    def __init__(self, device: torch.device) -> None:
        super().__init__()
        self._device = device

    def forward(self, tensors: List[torch.Tensor]) -> torch.Tensor:
        return torch.cat([t.to(self._device, non_blocking=True) for t in tensors], dim=1)


# ---- single-process multi-GPU gathers (fbgemm merge_pooled_embeddings / sum_reduce_to_one / all_to_one_device) ----
def all_to_one_device(inputs: List[torch.Tensor], target_device: torch.device) -> List[torch.Tensor]:
    """Copy every tensor to ``target_device`` (NVLink P2P, no NCCL). Each source device's copy is issued on that device's
    current stream and the target stream waits for all of them, so the copies of different sources overlap."""
    target_device = torch.device(target_device)
    if target_device.type != "cuda":
        return [t.to(target_device) for t in inputs]
    out: List[torch.Tensor] = []
    events = []
    for t in inputs:
        if t.device == target_device:
            out.append(t)
            continue
        if t.is_cuda:
            with torch.cuda.device(t.device):
                moved = t.to(target_device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(t.device))
            events.append(ev)
        else:
            moved = t.to(target_device, non_blocking=True)
        out.append(moved)
    cur = torch.cuda.current_stream(target_device)
    for ev in events:
        cur.wait_event(ev)
    return out


def merge_pooled_embeddings(pooled_embeddings: List[torch.Tensor], uncat_dim_size: int, target_device: torch.device, cat_dim: int = 1) -> torch.Tensor:
    """Gather per-device pooled embeddings onto ``target_device`` and concatenate along ``cat_dim`` (the other dim has
    ``uncat_dim_size`` entries)."""
    moved = all_to_one_device(pooled_embeddings, target_device)
    for t in moved:
        assert t.shape[1 - cat_dim] == uncat_dim_size, (t.shape, uncat_dim_size, cat_dim)
    return torch.cat(moved, dim=cat_dim) if len(moved) > 1 else moved[0]


def sum_reduce_to_one(inputs: List[torch.Tensor], target_device: torch.device) -> torch.Tensor:
    """Sum same-shaped per-device tensors on ``target_device`` (row-wise sharded inference output)."""
    moved = all_to_one_device(inputs, target_device)
    if len(moved) == 1:
        return moved[0]
    return torch.stack(moved, dim=0).sum(dim=0)


class EmbeddingsAllToOne(nn.Module):
    def __init__(self, device: torch.device, world_size: int, cat_dim: int) -> None:
        super().__init__()
        self._device = device
        self._world_size = world_size
        self._cat_dim = cat_dim

    def forward(self, tensors: List[torch.Tensor]) -> torch.Tensor:
        assert len(tensors) <= self._world_size
        return merge_pooled_embeddings(tensors, tensors[0].shape[1 - self._cat_dim], self._device, self._cat_dim)


class EmbeddingsAllToOneReduce(nn.Module):
    def __init__(self, device: torch.device, world_size: int) -> None:
        super().__init__()
        self._device = device
        self._world_size = world_size

    def forward(self, tensors: List[torch.Tensor]) -> torch.Tensor:
        return sum_reduce_to_one(tensors, self._device)


class SeqEmbeddingsAllToOne(nn.Module):
    def __init__(self, device: torch.device, world_size: int) -> None:
        super().__init__()
        self._device = device
        self._world_size = world_size

    def forward(self, tensors: List[torch.Tensor]) -> List[torch.Tensor]:
        return all_to_one_device(tensors, self._device)
