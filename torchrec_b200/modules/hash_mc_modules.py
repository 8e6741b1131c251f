"""Multi-probe zero-collision hash (MPZCH) managed collision module
(reference torchrec/modules/hash_mc_modules.py:196, hash_mc_evictions.py, hash_mc_metrics.py).

Open-addressing table of ``zch_size`` slots: an id hashes to a start slot inside its bucket and probes up to
``max_probe`` consecutive slots for itself or for a free / evictable slot. Eviction scoring is pluggable (none, LRU by
last-seen hour, TTL). Replaces ``fbgemm.zero_collision_hash`` / ``create_zch_buffer``. On CUDA the whole probe is ONE kernel
(``ops/csrc/zch.cu``: thread per id, atomicCAS slot claims, no host round trip, device-side statistics); on CPU a vectorised
PyTorch mirror with the same hash, so tables built on either side are interchangeable."""
from __future__ import annotations

import logging
import time
from enum import Enum, unique
from typing import Any, Dict, List, Optional, Tuple

import torch

from ..ops import _lib
from ..sparse.jagged_tensor import JaggedTensor
from .mc_modules import ManagedCollisionModule

logger = logging.getLogger(__name__)


@unique
class HashZchEvictionPolicyName(Enum):
    SINGLE_TTL_EVICTION = "SINGLE_TTL_EVICTION"
    LRU_EVICTION = "LRU_EVICTION"
    NONE = "NONE"


class HashZchEvictionConfig:
    def __init__(self, features: List[str], single_ttl: Optional[int] = None) -> None:
        self.features = features
        self.single_ttl = single_ttl


class ScalarLogger(torch.nn.Module):
    """Running hit / insert / collision statistics of one ZCH table (reference hash_mc_metrics.py)."""

    def __init__(self, name: str, zch_size: int, frequency: int = 100, start_bucket: int = 0, log_file_path: str = "") -> None:
        super().__init__()
        self._name, self._zch_size, self._frequency = name, zch_size, frequency
        self._steps = 0
        self.stats: Dict[str, float] = {"hit": 0, "insert": 0, "collision": 0, "total": 0}

    def update(self, hit: int, insert: int, collision: int, total: int) -> None:
        self.stats["hit"] += hit
        self.stats["insert"] += insert
        self.stats["collision"] += collision
        self.stats["total"] += total
        self._steps += 1
        if self._steps % self._frequency == 0:
            t = max(self.stats["total"], 1)
            logger.info(f"{self._name}: hit rate {self.stats['hit'] / t:.4f} insert rate {self.stats['insert'] / t:.4f} collision rate {self.stats['collision'] / t:.4f}")

    def forward(self) -> Dict[str, float]:
        return dict(self.stats)


def _mix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64-style finaliser on int64 tensors (wrap-around arithmetic)."""
    x = x ^ (x >> 30)
    x = x * -4658895280553007687  # 0xbf58476d1ce4e5b9
    x = x ^ (x >> 27)
    x = x * -7723592293110705685  # 0x94d049bb133111eb
    return x ^ (x >> 31)


class TrainInputMapper(torch.nn.Module):
    """Maps inference ids to the chunk of the identity table they lived in during training. Row-wise sharded training gives every
    dispatch unit (a training rank, or one of its buckets) its own range of the identity table; for serving the ranges are merged
    into one table and every id must be probed inside the range its training unit owned (reference hash_mc_modules.py:82).

    ``forward(values, output_offset)`` -> (values, local_sizes, offsets): per id the size and the start of its unit's range
    (minus ``output_offset``). The unit of an id follows the input dist of training: ``"modulo"`` (``id % units`` - the reference's rule
    for ``input_hash_size == 0``), ``"block"`` (``id // ceil(input_hash_size / units)`` - the reference's rule otherwise) or ``"hash"``
    (the multiplicative hash this framework's sharded managed-collision collection dispatches with); ``"auto"`` picks modulo / block
    from ``input_hash_size`` like the reference. With ``inference_dispatch_div_train_world_size`` the ids are reduced the way the
    training input dist reduced them (``id // units`` / ``id % block``)."""

    def __init__(self, input_hash_size: int, total_num_buckets: int, size_per_rank: torch.Tensor, train_rank_offsets: torch.Tensor,
                 inference_dispatch_div_train_world_size: bool = False, name: Optional[str] = None, dispatch: str = "auto") -> None:
        super().__init__()
        assert total_num_buckets > 0, f"{total_num_buckets=} must be positive"
        assert dispatch in ("auto", "modulo", "block", "hash")
        self._input_hash_size = input_hash_size
        self._buckets = total_num_buckets
        self._inference_dispatch_div_train_world_size = inference_dispatch_div_train_world_size
        self._name = name
        self._dispatch = dispatch if dispatch != "auto" else ("modulo" if input_hash_size == 0 else "block")
        self.register_buffer("_zch_size_per_training_rank", size_per_rank.to(torch.int64), persistent=False)
        self.register_buffer("_train_rank_offsets", train_rank_offsets.to(torch.int64), persistent=False)

    def _get_values_sizes_offsets(self, x: torch.Tensor, output_offset: Optional[torch.Tensor]):
        sizes, offs = self._zch_size_per_training_rank.to(x.device), self._train_rank_offsets.to(x.device)
        if self._dispatch == "hash":
            h = (x ^ (x >> 31)) * -7046029254386353131
            unit = torch.remainder(h ^ (h >> 29), self._buckets)
        elif self._dispatch == "modulo":
            unit = x % self._buckets
            if self._inference_dispatch_div_train_world_size:
                x = x // self._buckets
        else:
            blk = self._input_hash_size // self._buckets + (0 if self._input_hash_size % self._buckets == 0 else 1)
            unit = (x // blk).clamp(max=self._buckets - 1)
            if self._inference_dispatch_div_train_world_size:
                x = x % blk
        local_sizes = sizes.index_select(0, unit)
        offsets = offs.index_select(0, unit)
        if output_offset is not None:
            offsets = offsets - output_offset
        return x, local_sizes, offsets

    def forward(self, values: torch.Tensor, output_offset: Optional[torch.Tensor] = None):
        return self._get_values_sizes_offsets(values.to(torch.int64), output_offset)


class HashZchManagedCollisionModule(ManagedCollisionModule):
    def __init__(self, zch_size: int, device: torch.device, total_num_buckets: int, max_probe: int = 128, input_hash_size: int = (2**63) - 1,
                 output_segments: Optional[List[int]] = None, is_inference: bool = False, name: Optional[str] = None, tb_logging_frequency: int = 0,
                 eviction_policy_name: Optional[HashZchEvictionPolicyName] = None, eviction_config: Optional[HashZchEvictionConfig] = None,
                 inference_dispatch_div_train_world_size: bool = False, start_bucket: int = 0, end_bucket: Optional[int] = None,
                 opt_in_prob: int = -1, percent_reserved_slots: float = 0, disable_fallback: bool = False, track_id_freq: bool = False,
                 read_only_suffix: str = "_readonly", enable_per_feature_lookups: bool = False, no_bag: bool = False, write_runtime_meta_dim: int = 0) -> None:
        if output_segments is None:
            assert zch_size % total_num_buckets == 0, f"please pass output segments if not uniform buckets {zch_size=}, {total_num_buckets=}"
            output_segments = [(zch_size // total_num_buckets) * bucket for bucket in range(total_num_buckets + 1)]
        super().__init__(device=device, output_segments=output_segments, skip_state_validation=True, read_only_suffix=read_only_suffix,
                         enable_per_feature_lookups=enable_per_feature_lookups)
        self._track_id_freq, self._no_bag, self._write_runtime_meta_dim = track_id_freq, no_bag, write_runtime_meta_dim  # recorded for tooling; the probe keeps one 32-bit metadata word per slot
        self._zch_size_total = zch_size
        self._total_num_buckets = total_num_buckets
        self._start_bucket = start_bucket
        self._end_bucket = end_bucket if end_bucket is not None else total_num_buckets
        self._output_global_offset_tensor: Optional[torch.Tensor] = None
        self._name = name
        self._is_inference = is_inference
        self._max_probe = max_probe
        self._input_hash_size = input_hash_size
        self._eviction_policy_name = eviction_policy_name or HashZchEvictionPolicyName.NONE
        self._eviction_config = eviction_config
        self._disable_fallback = disable_fallback
        lo, hi = output_segments[self._start_bucket], output_segments[self._end_bucket]
        self._zch_size = hi - lo
        self._offset = lo
        self._buckets_local = self._end_bucket - self._start_bucket
        self._bucket_size = self._zch_size // max(self._buckets_local, 1)
        self.register_buffer("_hash_zch_identities", torch.full((self._zch_size, 1), -1, dtype=torch.int64, device=device))
        self.register_buffer("_hash_zch_metadata", torch.zeros((self._zch_size, 1), dtype=torch.int32, device=device))
        self._scalar_logger = ScalarLogger(name or "zch", self._zch_size, tb_logging_frequency) if tb_logging_frequency > 0 else None
        self._evicted_indices: List[torch.Tensor] = []
        self._evicted_pending: List[torch.Tensor] = []  # CUDA probe: per-id evicted slot or -1, compacted when evict() is called
        self._counters: Optional[torch.Tensor] = None
        self._device_steps = 0
        self._train_input_mapper: Optional[TrainInputMapper] = None
        self._train_unit_buckets = 1

    def preprocess(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        return features

    def _now(self) -> int:
        return int(time.time() // 3600)

    _POLICY_CODE = {HashZchEvictionPolicyName.NONE: 0, HashZchEvictionPolicyName.SINGLE_TTL_EVICTION: 1, HashZchEvictionPolicyName.LRU_EVICTION: 2}

    @torch.no_grad()
    def _probe_cuda(self, ids: torch.Tensor, readonly: bool) -> torch.Tensor:
        """One launch of ``trb_zch_probe`` for the whole id list; nothing is read back on the host."""
        import ctypes

        ids = ids.contiguous()
        out = torch.empty_like(ids)
        policy = self._POLICY_CODE[self._eviction_policy_name]
        evicted = torch.empty_like(ids) if (not readonly and policy != 0) else None
        if self._counters is None or self._counters.device != ids.device:
            self._counters = torch.zeros(4, dtype=torch.int32, device=ids.device)
        ttl = self._eviction_config.single_ttl if (self._eviction_config and self._eviction_config.single_ttl) else -1
        code = _lib.lib().trb_zch_probe(_lib.ptr(ids), ctypes.c_int64(ids.numel()), _lib.ptr(self._hash_zch_identities), _lib.ptr(self._hash_zch_metadata),
                                        ctypes.c_int64(max(self._buckets_local, 1)), ctypes.c_int64(max(self._bucket_size, 1)), int(self._max_probe), int(readonly),
                                        int(self._now()), int(ttl), policy, int(not self._disable_fallback), _lib.ptr(out), _lib.ptr(evicted),
                                        _lib.ptr(self._counters), _lib.stream_ptr(ids.device))
        _lib.check(code, "trb_zch_probe")
        if evicted is not None:
            self._evicted_pending.append(evicted)
        if self._scalar_logger is not None:
            self._device_steps += 1
            if self._device_steps % max(self._scalar_logger._frequency, 1) == 0:  # the only host read, once per logging period
                self.flush_statistics()
        return out

    def flush_statistics(self) -> Dict[str, float]:
        """Fold the device-side counters of the CUDA probe into the scalar logger (synchronises)."""
        if self._counters is not None:
            h, i, c, _ = (int(v) for v in self._counters.tolist())
            self._counters.zero_()
            if self._scalar_logger is not None and (h or i or c):
                self._scalar_logger.update(h, i, c, h + i + c)
        return self._scalar_logger() if self._scalar_logger is not None else {}

    @torch.no_grad()
    def set_train_input_mapper(self, mapper: Optional[TrainInputMapper], buckets_per_unit: int = 1) -> None:
        """Serving a table whose identities were trained row-wise sharded and merged afterwards: every id is probed inside the range of
        its training unit (``buckets_per_unit`` buckets each, as the unit's module had them). Lookups only."""
        self._train_input_mapper = mapper
        self._train_unit_buckets = buckets_per_unit

    @classmethod
    def merge_trained_shards(cls, shards: List["HashZchManagedCollisionModule"], device: Optional[torch.device] = None, **kwargs) -> "HashZchManagedCollisionModule":
        """One serving module from the per-rank modules of a row-wise sharded training run (rank order): identities / metadata are
        concatenated, a ``TrainInputMapper`` with this framework's dispatch hash sends every id to its training rank's range."""
        first = shards[0]
        dev = device or first._hash_zch_identities.device
        sizes = [m._zch_size for m in shards]
        total = sum(sizes)
        merged = cls(zch_size=total, device=dev, total_num_buckets=len(shards), max_probe=first._max_probe, input_hash_size=first._input_hash_size,
                     output_segments=[sum(sizes[:i]) for i in range(len(shards) + 1)], is_inference=True, name=first._name, disable_fallback=first._disable_fallback, **kwargs)
        merged._hash_zch_identities.copy_(torch.cat([m._hash_zch_identities.to(dev) for m in shards]))
        merged._hash_zch_metadata.copy_(torch.cat([m._hash_zch_metadata.to(dev) for m in shards]))
        offs = torch.tensor([sum(sizes[:i]) for i in range(len(shards))], dtype=torch.int64)
        merged.set_train_input_mapper(TrainInputMapper(first._input_hash_size, len(shards), torch.tensor(sizes, dtype=torch.int64), offs, dispatch="hash", name=first._name),
                                      buckets_per_unit=first._buckets_local)
        return merged

    def _probe_mapped(self, ids: torch.Tensor) -> torch.Tensor:
        vals, sizes, offsets = self._train_input_mapper(ids)
        out = torch.empty_like(ids)
        ident, meta = self._hash_zch_identities.view(-1), self._hash_zch_metadata.view(-1)
        for off in torch.unique(offsets).tolist():
            m = offsets == off
            size = int(sizes[m][0])
            nb = max(self._train_unit_buckets, 1)
            out[m] = self._probe_torch(vals[m], True, ident[off : off + size], meta[off : off + size], nb, size // nb) + off
        return out

    def _probe(self, ids: torch.Tensor, readonly: bool) -> torch.Tensor:
        n = ids.numel()
        if n == 0:
            return ids
        if self._train_input_mapper is not None:
            return self._probe_mapped(ids)
        if _lib.use_cuda_kernels(ids):
            return self._probe_cuda(ids, readonly)
        return self._probe_torch(ids, readonly, self._hash_zch_identities.view(-1), self._hash_zch_metadata.view(-1), self._buckets_local, self._bucket_size)

    def _probe_torch(self, ids: torch.Tensor, readonly: bool, ident: torch.Tensor, meta: torch.Tensor, buckets_local: int, bucket_size: int) -> torch.Tensor:
        n = ids.numel()
        h = _mix64(ids)
        bucket = torch.remainder(h, max(buckets_local, 1))
        start = torch.remainder(h >> 16, max(bucket_size, 1))
        out = torch.full_like(ids, -1)
        pending = torch.ones(n, dtype=torch.bool, device=ids.device)
        now = self._now()
        ttl = self._eviction_config.single_ttl if (self._eviction_config and self._eviction_config.single_ttl) else None
        hits = inserts = 0
        for p in range(min(self._max_probe, max(bucket_size, 1))):
            if not bool(pending.any()):
                break
            idx = pending.nonzero(as_tuple=True)[0]
            slot = bucket[idx] * bucket_size + torch.remainder(start[idx] + p, max(bucket_size, 1))
            cur = ident[slot]
            hit = cur == ids[idx]
            out[idx[hit]] = slot[hit]
            pending[idx[hit]] = False
            hits += int(hit.sum())
            if readonly:
                continue
            free = cur == -1
            if self._eviction_policy_name == HashZchEvictionPolicyName.SINGLE_TTL_EVICTION and ttl is not None:
                free = free | ((cur != -1) & ~hit & (meta[slot].long() + ttl < now))
            cand = idx[free & ~hit]
            if cand.numel():
                cslot = slot[free & ~hit]
                # several ids may want the same free slot in this round: the first one wins, the rest keep probing
                uniq_slot, first = _first_occurrence(cslot)
                # also de-duplicate identical ids (they will hit on the next probe round)
                win_ids = ids[cand[first]]
                old = ident[uniq_slot]
                if bool((old != -1).any()):
                    self._evicted_indices.append(uniq_slot[old != -1] + self._offset)
                ident[uniq_slot] = win_ids
                meta[uniq_slot] = now
                out[cand[first]] = uniq_slot
                pending[cand[first]] = False
                inserts += int(uniq_slot.numel())
                # duplicates of a winning id: resolve immediately
                same = pending.clone()
                same[:] = False
                rest = cand[~_mask_of(first, cand.numel(), cand.device)]
                if rest.numel():
                    rs = cslot[~_mask_of(first, cand.numel(), cand.device)]
                    again = ident[rs] == ids[rest]
                    out[rest[again]] = rs[again]
                    pending[rest[again]] = False
        touched = out >= 0
        if not readonly and bool(touched.any()):
            meta[out[touched]] = now
        if self._scalar_logger is not None:
            self._scalar_logger.update(hits, inserts, int(pending.sum()), n)
        # ids that found no slot fall back to their start slot (collision) unless disabled
        fb = bucket * bucket_size + start
        out = torch.where(out >= 0, out, fb if not self._disable_fallback else torch.full_like(out, -1))
        return out

    def remap(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        res: Dict[str, JaggedTensor] = {}
        readonly = self._is_inference or not self.training
        for name, f in features.items():
            slot = self._probe(f.values().to(torch.int64), readonly)
            res[name] = JaggedTensor(values=slot + self._offset, lengths=f.lengths(), offsets=f.offsets(), weights=f.weights_or_none())
        return res

    def profile(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        return features

    def forward(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        return self.remap(features)

    def output_size(self) -> int:
        return self._zch_size_total

    def buckets(self) -> int:
        return self._total_num_buckets

    def input_size(self) -> int:
        return self._input_hash_size

    def open_slots(self) -> torch.Tensor:
        return (self._hash_zch_identities.view(-1) == -1).sum().view(1)

    def evict(self) -> Optional[torch.Tensor]:
        for t in self._evicted_pending:
            t = t[t >= 0]
            if t.numel():
                self._evicted_indices.append(t + self._offset)
        self._evicted_pending = []
        if not self._evicted_indices:
            return None
        out = torch.cat(self._evicted_indices)
        self._evicted_indices = []
        return out

    def rebuild_with_output_id_range(self, output_id_range: Tuple[int, int], output_segments: List[int], device: Optional[torch.device] = None):
        start_bucket = output_segments.index(output_id_range[0])
        end_bucket = output_segments.index(output_id_range[1])
        return type(self)(zch_size=self._zch_size_total, device=device or self.device, total_num_buckets=self._total_num_buckets, max_probe=self._max_probe,
                          input_hash_size=self._input_hash_size, output_segments=output_segments, is_inference=self._is_inference, name=self._name,
                          eviction_policy_name=self._eviction_policy_name, eviction_config=self._eviction_config, start_bucket=start_bucket, end_bucket=end_bucket,
                          disable_fallback=self._disable_fallback)


def _first_occurrence(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Unique values of x and the index of their first occurrence."""
    uniq, inv = torch.unique(x, return_inverse=True)
    first = torch.full((uniq.numel(),), x.numel(), dtype=torch.long, device=x.device)
    first.scatter_reduce_(0, inv, torch.arange(x.numel(), device=x.device), reduce="amin", include_self=True)
    return uniq, first


def _mask_of(idx: torch.Tensor, n: int, device) -> torch.Tensor:
    m = torch.zeros(n, dtype=torch.bool, device=device)
    m[idx] = True
    return m
