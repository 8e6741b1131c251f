"""Shard estimators under their reference import path, plus the functional storage helpers.

Reference: ``torchrec/distributed/planner/shard_estimators.py`` - ``EmbeddingPerfEstimator`` :70 (legacy, monolithic), ``EmbeddingStorageEstimator`` :117,
``calculate_pipeline_io_cost`` :260, ``calculate_shard_storages`` :303, ``get_num_poolings`` :495, ``EmbeddingOffloadStats`` :889.
The estimator classes are implemented in ``enumerators.py`` (the annotation-driven perf estimator in ``estimator/``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from ..embedding_types import EmbeddingComputeKernel
from ..types import CacheStatistics, PipelineType, ShardingType
from .constants import BIGINT_DTYPE, UVM_CACHING_RATIO
from .enumerators import EmbeddingPerfEstimator, EmbeddingStorageEstimator  # noqa: F401
from .types import ParameterConstraints, ShardingOption, Storage, Topology


def get_num_poolings(constraints: Optional[Dict[str, ParameterConstraints]], so: ShardingOption) -> List[float]:
    c = constraints.get(so.name) if constraints else None
    if c is not None and c.num_poolings:
        assert len(c.num_poolings) == so.num_inputs, f"{so.name}: {len(c.num_poolings)} num_poolings for {so.num_inputs} inputs"
        return list(c.num_poolings)
    return [1.0] * so.num_inputs


def calculate_pipeline_io_cost(input_size: int, output_size: int, prefetch_size: int, pipeline_type: PipelineType, multipass_prefetch_max_pass: Optional[int],
                               count_ephemeral_storage_cost: bool = False, is_inference: bool = False) -> int:
    """HBM held by in-flight batches. The sparse-dist pipeline keeps the current AND the next batch's distributed ids (x2); the prefetch pipeline one
    more (x3) plus the cache-fill burst, which multi-pass prefetch divides by its pass count. Outputs live only across the all-to-all: counted only
    when the caller wants ephemeral peaks."""
    if is_inference:
        return 0
    out = output_size if count_ephemeral_storage_cost else 0
    if pipeline_type == PipelineType.TRAIN_SPARSE_DIST:
        return 2 * input_size + out
    if pipeline_type == PipelineType.TRAIN_PREFETCH_SPARSE_DIST:
        passes = multipass_prefetch_max_pass or 1
        return 3 * input_size + int((1 + 6 / passes) * prefetch_size) + out
    return input_size + output_size


def _io_sizes(sharding_type: str, batch_sizes: List[int], world_size: int, local_world_size: int, input_lengths: List[float], emb_dim: int,
              shard_sizes: List[List[int]], input_data_type_size: float, output_data_type_size: float, num_poolings: List[float], is_pooled: bool) -> Tuple[List[int], List[int]]:
    """(input bytes, output bytes) per shard: what arrives at / leaves the rank holding it per step."""
    ids_local = sum(l * p * b for l, p, b in zip(input_lengths, num_poolings, batch_sizes))
    bags_local = sum(p * b for p, b in zip(num_poolings, batch_sizes))
    ins, outs = [], []
    n_row_shards = len({s[0] for s in shard_sizes}) if sharding_type == ShardingType.ROW_WISE.value else 1
    for rows, cols in shard_sizes:
        if sharding_type == ShardingType.DATA_PARALLEL.value:
            ids, bags = ids_local, bags_local
        elif sharding_type == ShardingType.ROW_WISE.value:
            ids, bags = ids_local * world_size / max(world_size, 1), bags_local * world_size
        elif sharding_type in (ShardingType.TABLE_ROW_WISE.value, ShardingType.GRID_SHARD.value):
            ids, bags = ids_local * world_size / max(local_world_size, 1), bags_local * world_size
        else:
            ids, bags = ids_local * world_size, bags_local * world_size
        ins.append(int(math.ceil(ids * input_data_type_size)))
        outs.append(int(math.ceil((bags if is_pooled else ids) * cols * output_data_type_size)))
    return ins, outs


def calculate_shard_storages(sharder, sharding_type: str, tensor: torch.Tensor, compute_device: str, compute_kernel: str, shard_sizes: List[List[int]],
                             batch_sizes: List[int], world_size: int, local_world_size: int, input_lengths: List[float], num_poolings: List[float],
                             caching_ratio: float, is_pooled: bool, input_data_type_size: float = BIGINT_DTYPE, output_data_type_size: float = 4.0,
                             pipeline_type: PipelineType = PipelineType.NONE, count_ephemeral_storage_cost: bool = False, is_inference: bool = False,
                             multipass_prefetch_max_pass: Optional[int] = None, key_value_params=None, optimizer_multiplier: Optional[float] = None) -> List[Storage]:
    """Storage of every shard of one sharding option: weights + optimizer state (by optimizer class on the tensor) + pipeline IO."""
    elem = tensor.element_size()
    ins, outs = _io_sizes(sharding_type, batch_sizes, world_size, local_world_size, input_lengths, tensor.shape[1], shard_sizes, input_data_type_size,
                          output_data_type_size, num_poolings, is_pooled)
    if optimizer_multiplier is None:
        names = {c.__name__ for c in (getattr(tensor, "_optimizer_classes", None) or [])}
        if is_inference or compute_kernel.startswith("quant"):
            optimizer_multiplier = 0.0
        elif names & {"Adam", "AdamW", "LAMB"}:
            optimizer_multiplier = 2.0
        elif names & {"Adagrad", "PartialRowWiseAdam", "PartialRowWiseLAMB"}:
            optimizer_multiplier = 1.0
        elif names & {"SGD", "LarsSGD"}:
            optimizer_multiplier = 0.0
        else:
            optimizer_multiplier = -1.0  # row-wise: one fp32 per row
    out: List[Storage] = []
    cached = compute_kernel in (EmbeddingComputeKernel.FUSED_UVM_CACHING.value, EmbeddingComputeKernel.QUANT_UVM_CACHING.value)
    host = compute_kernel in (EmbeddingComputeKernel.FUSED_UVM.value, EmbeddingComputeKernel.QUANT_UVM.value, EmbeddingComputeKernel.KEY_VALUE.value)
    for (rows, cols), i, o in zip(shard_sizes, ins, outs):
        w = rows * cols * elem
        opt = rows * 4 if optimizer_multiplier < 0 else rows * cols * 4 * optimizer_multiplier
        hbm = ddr = 0.0
        if host:
            ddr = w + opt
        elif cached:
            ratio = caching_ratio if caching_ratio is not None else UVM_CACHING_RATIO
            ddr = w + opt
            hbm = ratio * (w + opt) + rows * 4  # + cache aux state (slot map)
        elif compute_device == "cuda":
            hbm = w + opt
        else:
            ddr = w + opt
        prefetch_size = i if (cached and pipeline_type == PipelineType.TRAIN_PREFETCH_SPARSE_DIST) else 0
        io = calculate_pipeline_io_cost(i, o, prefetch_size, pipeline_type, multipass_prefetch_max_pass, count_ephemeral_storage_cost, is_inference)
        if compute_device == "cuda":
            hbm += io
        else:
            ddr += io
        out.append(Storage(hbm=int(math.ceil(hbm)), ddr=int(math.ceil(ddr))))
    return out


class EmbeddingOffloadStats(CacheStatistics):
    """Cache statistics of a host-offloaded table from its miss-rate curve (MRC).

    ``mrc_hist_counts[i]`` = number of lookups whose reuse distance falls in bucket ``i`` of ``height`` (unique rows) evenly spaced buckets (+ one
    overflow bucket for cold misses). ``expected_miss_rate(clf)`` integrates the tail of the histogram above the cache size ``clf * height``;
    ``cacheability`` is the area under the miss-rate curve (0 = perfectly cacheable, 1 = every size misses everything)."""

    def __init__(self, cacheability: float, expected_lookups: int, mrc_hist_counts: torch.Tensor, height: int) -> None:
        self._cacheability = cacheability
        self._expected_lookups = expected_lookups
        self.height = height
        if mrc_hist_counts.dim() != 1:
            raise ValueError(f"expected 1d tensor, got {mrc_hist_counts.dim()}d")
        if mrc_hist_counts.numel() == 0:
            raise ValueError("expected non-empty tensor")
        self.hist = mrc_hist_counts
        self.bins = torch.linspace(0, height, mrc_hist_counts.numel())

    @property
    def expected_lookups(self) -> int:
        return self._expected_lookups

    def expected_miss_rate(self, clf: float) -> float:
        return float(EmbeddingOffloadStats.estimate_cache_miss_rate(torch.tensor([clf * self.height]), self.hist, self.bins)[0])

    @property
    def cacheability(self) -> float:
        return self._cacheability

    @staticmethod
    def estimate_cache_miss_rate(cache_sizes: torch.Tensor, hist: torch.Tensor, bins: torch.Tensor) -> torch.Tensor:
        """Miss rate at each cache size: share of lookups whose reuse distance is >= the size (linear inside a bucket)."""
        total = hist.sum()
        if float(total) == 0:
            return torch.zeros_like(cache_sizes, dtype=torch.float32)
        tail = torch.flip(torch.cumsum(torch.flip(hist.double(), [0]), 0), [0]) / total  # misses if cache < bins[i]
        idx = torch.bucketize(cache_sizes.double(), bins.double(), right=False).clamp(max=hist.numel() - 1)
        lo = (idx - 1).clamp(min=0)
        span = (bins[idx] - bins[lo]).double().clamp(min=1e-12)
        frac = ((cache_sizes.double() - bins[lo].double()) / span).clamp(0, 1)
        at_lo = torch.where(idx == 0, torch.ones_like(frac), tail[lo])
        return (at_lo + (tail[idx] - at_lo) * frac).float()
