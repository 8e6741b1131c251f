"""Managed-collision (zero-collision-hash) modules (reference torchrec/modules/mc_modules.py:175-1300).

A ``ManagedCollisionModule`` remaps raw (unbounded) ids to a bounded table range ``[0, zch_size)``: ids it has
admitted keep a stable slot; unseen ids compete for slots according to an eviction policy (LFU / LRU /
distance-LFU) every ``eviction_interval`` steps; until admitted they fall back to a hashed slot.
``MCHManagedCollisionModule`` is the sorted-table ("sorted ZCH") implementation: a sorted id vector + searchsorted.
"""
from __future__ import annotations

import abc
from dataclasses import dataclass
from typing import Callable, Dict, List, NamedTuple, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch import nn

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor


def apply_mc_method_to_jt_dict(mc_module: nn.Module, method: str, features_dict: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
    """``getattr(mc_module, method)(features_dict)`` - the named step (``preprocess`` / ``profile`` / ``remap``) of a managed-collision
    module on a dict of jagged tensors, same key order out (reference mc_modules.py:24)."""
    return getattr(mc_module, method)(features_dict)


class ManagedCollisionModule(nn.Module):
    """Abstract remapper of the ids of one table."""

    def __init__(self, device: torch.device, output_segments: Optional[List[int]] = None, skip_state_validation: bool = False, read_only_suffix: str = "_readonly",
                 enable_per_feature_lookups: bool = False) -> None:
        super().__init__()
        self._device = device
        self._output_segments = output_segments
        self._skip_state_validation = skip_state_validation
        self._read_only_suffix = read_only_suffix  # features named <feature><suffix> are looked up without being admitted (see ``readable_suffix``)
        self._enable_per_feature_lookups = enable_per_feature_lookups

    @property
    def readable_suffix(self) -> str:
        return self._read_only_suffix

    @abc.abstractmethod
    def preprocess(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        ...

    @property
    def device(self) -> torch.device:
        return self._device

    @abc.abstractmethod
    def evict(self) -> Optional[torch.Tensor]:
        """Slots whose ids were replaced since the last call (their embedding rows should be reset)."""

    @abc.abstractmethod
    def reset_inference_mode(self) -> None:
        """Switch to serving: lookups only (no admission, no eviction, no statistics)."""
        self.train(False)
        if hasattr(self, "_is_inference"):
            self._is_inference = True

    def remap(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        ...

    @abc.abstractmethod
    def profile(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        ...

    def forward(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        if getattr(self, "_need_preprocess", True):
            features = self.preprocess(features)
        if self.training:
            self.profile(features)
        return self.remap(features)

    @abc.abstractmethod
    def output_size(self) -> int:
        ...

    @abc.abstractmethod
    def input_size(self) -> int:
        ...

    def buckets(self) -> int:
        return 1

    def validate_state(self) -> None:
        pass

    @abc.abstractmethod
    def rebuild_with_output_id_range(self, output_id_range: Tuple[int, int], output_segments: List[int], device: Optional[torch.device] = None) -> "ManagedCollisionModule":
        ...

    @abc.abstractmethod
    def open_slots(self) -> torch.Tensor:
        ...


class MCHEvictionPolicyMetadataInfo(NamedTuple):
    metadata_name: str
    is_mch_metadata: bool
    is_history_metadata: bool


class MCHEvictionPolicy(abc.ABC):
    """Decides which ids own the ZCH slots. Metadata tensors ride along the sorted id table."""

    def __init__(self, metadata_info: List[MCHEvictionPolicyMetadataInfo], threshold_filtering_func: Optional[Callable[[torch.Tensor], Tuple[torch.Tensor, Union[float, torch.Tensor]]]] = None) -> None:
        self._metadata_info = metadata_info
        self._threshold_filtering_func = threshold_filtering_func

    @property
    @abc.abstractmethod
    def metadata_info(self) -> List[MCHEvictionPolicyMetadataInfo]:
        ...

    @abc.abstractmethod
    def record_history_metadata(self, current_iter: int, incoming_ids: torch.Tensor, history_metadata: Dict[str, torch.Tensor]) -> None:
        ...

    @abc.abstractmethod
    def coalesce_history_metadata(self, current_iter: int, history_metadata: Dict[str, torch.Tensor], unique_ids_counts: torch.Tensor,
                                  unique_inverse_mapping: torch.Tensor, additional_ids: Optional[torch.Tensor] = None,
                                  threshold_mask: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        ...

    @abc.abstractmethod
    def update_metadata_and_generate_eviction_scores(self, current_iter: int, mch_size: int, coalesced_history_argsort_mapping: torch.Tensor,
                                                     coalesced_history_sorted_unique_ids_counts: torch.Tensor, coalesced_history_mch_matching_elements_mask: torch.Tensor,
                                                     coalesced_history_mch_matching_indices: torch.Tensor, mch_metadata: Dict[str, torch.Tensor],
                                                     coalesced_history_metadata: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        ...

    def _compute_selected_eviction_and_replacement_indices(self, pivot: int, evict_scores: torch.Tensor, new_scores: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        # the `pivot` new ids with the highest scores replace the `pivot` resident ids with the lowest scores,
        # but only where the newcomer actually beats the resident
        n = min(pivot, new_scores.numel(), evict_scores.numel())
        if n == 0:
            e = torch.zeros(0, dtype=torch.long, device=evict_scores.device)
            return e, e
        new_sorted, new_idx = torch.sort(new_scores, descending=True, stable=True)
        ev_sorted, ev_idx = torch.sort(evict_scores, descending=False, stable=True)
        wins = new_sorted[:n] > ev_sorted[:n]
        return ev_idx[:n][wins], new_idx[:n][wins]


class LFU_EvictionPolicy(MCHEvictionPolicy):
    """Least frequently used: score = access count."""

    def __init__(self, threshold_filtering_func=None) -> None:
        super().__init__(metadata_info=[MCHEvictionPolicyMetadataInfo("counts", is_mch_metadata=True, is_history_metadata=False)],
                         threshold_filtering_func=threshold_filtering_func)

    @property
    def metadata_info(self):
        return self._metadata_info

    def record_history_metadata(self, current_iter, incoming_ids, history_metadata) -> None:
        pass

    def coalesce_history_metadata(self, current_iter, history_metadata, unique_ids_counts, unique_inverse_mapping, additional_ids=None, threshold_mask=None):
        return {}

    def update_metadata_and_generate_eviction_scores(self, current_iter, mch_size, coalesced_history_argsort_mapping, coalesced_history_sorted_unique_ids_counts,
                                                     coalesced_history_mch_matching_elements_mask, coalesced_history_mch_matching_indices, mch_metadata,
                                                     coalesced_history_metadata):
        mch_counts = mch_metadata["counts"]
        mch_counts[coalesced_history_mch_matching_indices] += coalesced_history_sorted_unique_ids_counts[coalesced_history_mch_matching_elements_mask]
        new_counts = coalesced_history_sorted_unique_ids_counts[~coalesced_history_mch_matching_elements_mask]
        return mch_counts.float(), new_counts.float()


class LRU_EvictionPolicy(MCHEvictionPolicy):
    """Least recently used: score = -(age); ``decay_exponent`` shapes the recency score."""

    def __init__(self, decay_exponent: float = 1.0, threshold_filtering_func=None) -> None:
        super().__init__(metadata_info=[MCHEvictionPolicyMetadataInfo("last_access_iter", is_mch_metadata=True, is_history_metadata=True)],
                         threshold_filtering_func=threshold_filtering_func)
        self._decay_exponent = decay_exponent

    @property
    def metadata_info(self):
        return self._metadata_info

    def record_history_metadata(self, current_iter, incoming_ids, history_metadata) -> None:
        history_metadata["last_access_iter"] = torch.full_like(incoming_ids, current_iter)

    def coalesce_history_metadata(self, current_iter, history_metadata, unique_ids_counts, unique_inverse_mapping, additional_ids=None, threshold_mask=None):
        last = history_metadata["last_access_iter"]
        if additional_ids is not None:
            last = torch.cat([last, torch.zeros_like(additional_ids)])
        out = torch.zeros(unique_ids_counts.numel(), dtype=last.dtype, device=last.device)
        out.scatter_reduce_(0, unique_inverse_mapping, last, reduce="amax", include_self=False)
        if threshold_mask is not None:
            out = out[threshold_mask]
        return {"last_access_iter": out}

    def update_metadata_and_generate_eviction_scores(self, current_iter, mch_size, coalesced_history_argsort_mapping, coalesced_history_sorted_unique_ids_counts,
                                                     coalesced_history_mch_matching_elements_mask, coalesced_history_mch_matching_indices, mch_metadata,
                                                     coalesced_history_metadata):
        hist = coalesced_history_metadata["last_access_iter"][coalesced_history_argsort_mapping]
        mch = mch_metadata["last_access_iter"]
        mch[coalesced_history_mch_matching_indices] = hist[coalesced_history_mch_matching_elements_mask]
        new = hist[~coalesced_history_mch_matching_elements_mask]
        score = lambda t: -torch.pow((current_iter - t + 1).float(), self._decay_exponent)
        return score(mch), score(new)


class DistanceLFU_EvictionPolicy(MCHEvictionPolicy):
    """count / (age ^ decay_exponent): frequent AND recent ids win."""

    def __init__(self, decay_exponent: float = 1.0, threshold_filtering_func=None) -> None:
        super().__init__(metadata_info=[MCHEvictionPolicyMetadataInfo("counts", True, False), MCHEvictionPolicyMetadataInfo("last_access_iter", True, True)],
                         threshold_filtering_func=threshold_filtering_func)
        self._decay_exponent = decay_exponent

    @property
    def metadata_info(self):
        return self._metadata_info

    def record_history_metadata(self, current_iter, incoming_ids, history_metadata) -> None:
        history_metadata["last_access_iter"] = torch.full_like(incoming_ids, current_iter)

    def coalesce_history_metadata(self, current_iter, history_metadata, unique_ids_counts, unique_inverse_mapping, additional_ids=None, threshold_mask=None):
        last = history_metadata["last_access_iter"]
        if additional_ids is not None:
            last = torch.cat([last, torch.zeros_like(additional_ids)])
        out = torch.zeros(unique_ids_counts.numel(), dtype=last.dtype, device=last.device)
        out.scatter_reduce_(0, unique_inverse_mapping, last, reduce="amax", include_self=False)
        if threshold_mask is not None:
            out = out[threshold_mask]
        return {"last_access_iter": out}

    def update_metadata_and_generate_eviction_scores(self, current_iter, mch_size, coalesced_history_argsort_mapping, coalesced_history_sorted_unique_ids_counts,
                                                     coalesced_history_mch_matching_elements_mask, coalesced_history_mch_matching_indices, mch_metadata,
                                                     coalesced_history_metadata):
        hist_last = coalesced_history_metadata["last_access_iter"][coalesced_history_argsort_mapping]
        m = coalesced_history_mch_matching_elements_mask
        mch_metadata["counts"][coalesced_history_mch_matching_indices] += coalesced_history_sorted_unique_ids_counts[m]
        mch_metadata["last_access_iter"][coalesced_history_mch_matching_indices] = hist_last[m]
        score = lambda c, t: c.float() / torch.pow((current_iter - t + 1).float(), self._decay_exponent)
        return score(mch_metadata["counts"], mch_metadata["last_access_iter"]), score(coalesced_history_sorted_unique_ids_counts[~m], hist_last[~m])


def dynamic_threshold_filter(id_counts: torch.Tensor, threshold_skew_multiplier: float = 10.0) -> Tuple[torch.Tensor, torch.Tensor]:
    num_ids = id_counts.numel()
    total_count = id_counts.sum()
    BASE_THRESHOLD = 1 / num_ids
    threshold_mass = BASE_THRESHOLD * threshold_skew_multiplier
    threshold = threshold_mass * total_count
    return id_counts > threshold, threshold


def average_threshold_filter(id_counts: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    if id_counts.dtype != torch.float:
        id_counts = id_counts.float()
    threshold = id_counts.mean()
    return id_counts > threshold, threshold


def probabilistic_threshold_filter(id_counts: torch.Tensor, per_id_probability: float = 0.01) -> Tuple[torch.Tensor, torch.Tensor]:
    probability = torch.full_like(id_counts, 1 - per_id_probability, dtype=torch.float)
    id_probability = 1 - torch.pow(probability, id_counts)
    threshold = torch.rand(id_counts.size(), device=id_counts.device)
    return id_probability > threshold, threshold


class MCHManagedCollisionModule(ManagedCollisionModule):
    """Sorted-ZCH: slots [0, zch_size) hold the admitted ids in sorted order (searchsorted remap);
    ids that are not admitted hash into the residual range or, without one, into the whole table."""

    def __init__(self, zch_size: int, device: torch.device, eviction_policy: MCHEvictionPolicy, eviction_interval: int, input_hash_size: int = (2**63) - 1,
                 input_hash_func: Optional[Callable[[torch.Tensor, int], torch.Tensor]] = None, mch_size: Optional[int] = None,
                 mch_hash_func: Optional[Callable[[torch.Tensor, int], torch.Tensor]] = None, name: Optional[str] = None,
                 output_global_offset: int = 0, output_segments: Optional[List[int]] = None, buckets: int = 1) -> None:
        if output_segments is None:
            output_segments = [output_global_offset, output_global_offset + zch_size]
        super().__init__(device=device, output_segments=output_segments)
        self._name = name
        self._input_history_buffer_size: int = -1
        self._input_hash_size = input_hash_size
        self._zch_size: int = zch_size
        assert self._zch_size > 0, "zch_size must be > 0"
        self._mch_size: int = mch_size if mch_size is not None else 0  # residual (hashed) range at the end of the table
        self._output_global_offset: int = output_global_offset
        self._mch_hash_func = mch_hash_func
        self._input_hash_func = input_hash_func
        self._eviction_interval = eviction_interval
        assert self._eviction_interval > 0, "eviction_interval must be > 1"
        self._eviction_policy = eviction_policy
        self._current_iter: int = -1
        self._buckets = buckets
        self._init_buffers()
        self._mch_metadata: Dict[str, torch.Tensor] = {}
        self._history_metadata: Dict[str, torch.Tensor] = {}
        self._init_metadata_buffers()
        self._current_history_buffer_offset: int = 0
        self._evicted: bool = False
        self._last_eviction_iter: int = -1

    def _init_buffers(self) -> None:
        n = self._zch_size - self._mch_size
        self._n_slots = n
        self.register_buffer("_mch_sorted_raw_ids", torch.full((n,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=self.device))
        self.register_buffer("_mch_slots", torch.tensor([n], dtype=torch.int64, device=self.device), persistent=False)
        self.register_buffer("_delimiter", torch.tensor([torch.iinfo(torch.int64).max], dtype=torch.int64, device=self.device), persistent=False)
        self.register_buffer("_mch_remapped_ids_mapping", torch.arange(n, dtype=torch.int64, device=self.device))
        self.register_buffer("_evicted_emb_indices", torch.empty((1,), device=self.device), persistent=False)
        self._evicted_emb_indices = torch.empty((1,), device=self.device)

    def _init_metadata_buffers(self) -> None:
        for info in self._eviction_policy.metadata_info:
            if info.is_mch_metadata:
                name = f"_mch_{info.metadata_name}"
                self.register_buffer(name, torch.zeros(self._n_slots, dtype=torch.int64, device=self.device))
                self._mch_metadata[info.metadata_name] = getattr(self, name)

    def _init_history_buffers(self, features: Dict[str, JaggedTensor]) -> None:
        input_batch_value_size_cumsum = sum(f.values().numel() for f in features.values())
        self._input_history_buffer_size = int(input_batch_value_size_cumsum * self._eviction_interval * 1.25) + 1
        self._history_ids = torch.empty(self._input_history_buffer_size, dtype=torch.int64, device=self.device)
        self._history_extra: Dict[str, torch.Tensor] = {i.metadata_name: torch.empty(self._input_history_buffer_size, dtype=torch.int64, device=self.device)
                                                        for i in self._eviction_policy.metadata_info if i.is_history_metadata}

    def preprocess(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        if self._input_hash_func is None:
            return features
        return {k: JaggedTensor(values=self._input_hash_func(f.values(), self._input_hash_size), lengths=f.lengths(), offsets=f.offsets(), weights=f.weights_or_none())
                for k, f in features.items()}

    @torch.no_grad()
    def _coalesce_history(self) -> None:
        n = self._current_history_buffer_offset
        current_iter = self._current_iter
        ids = self._history_ids[:n]
        hist_meta = {k: v[:n] for k, v in self._history_extra.items()}
        uniq, inverse, counts = torch.unique(ids, return_inverse=True, return_counts=True)
        mask = None
        if self._eviction_policy._threshold_filtering_func is not None:
            mask, _ = self._eviction_policy._threshold_filtering_func(counts)
        coalesced_meta = self._eviction_policy.coalesce_history_metadata(current_iter, hist_meta, counts, inverse, threshold_mask=mask)
        if mask is not None:
            uniq, counts = uniq[mask], counts[mask]
        # uniq is sorted already
        argsort = torch.arange(uniq.numel(), device=uniq.device)
        sorted_ids = self._mch_sorted_raw_ids
        pos = torch.searchsorted(sorted_ids, uniq).clamp(max=self._n_slots - 1)
        matching = sorted_ids[pos] == uniq
        evict_scores, new_scores = self._eviction_policy.update_metadata_and_generate_eviction_scores(
            current_iter, self._n_slots, argsort, counts, matching, pos[matching], self._mch_metadata, coalesced_meta)
        new_ids = uniq[~matching]
        # free slots (never used) count as the weakest residents
        evict_scores = torch.where(sorted_ids == torch.iinfo(torch.int64).max, torch.full_like(evict_scores, float("-inf")), evict_scores)
        ev_idx, new_idx = self._eviction_policy._compute_selected_eviction_and_replacement_indices(min(new_ids.numel(), self._n_slots), evict_scores, new_scores)
        if ev_idx.numel() > 0:
            self._mch_sorted_raw_ids[ev_idx] = new_ids[new_idx]
            for name, meta in self._mch_metadata.items():
                if name == "counts":
                    meta[ev_idx] = counts[~matching][new_idx]
                elif name in coalesced_meta:
                    meta[ev_idx] = coalesced_meta[name][~matching][new_idx]
            evicted_slots = self._mch_remapped_ids_mapping[ev_idx]
            # keep the id vector sorted; the slot mapping travels with its id
            order = torch.argsort(self._mch_sorted_raw_ids, stable=True)
            self._mch_sorted_raw_ids.copy_(self._mch_sorted_raw_ids[order])
            self._mch_remapped_ids_mapping.copy_(self._mch_remapped_ids_mapping[order])
            for meta in self._mch_metadata.values():
                meta.copy_(meta[order])
            self._evicted_emb_indices = evicted_slots + self._output_global_offset
            self._evicted = True
        self._current_history_buffer_offset = 0

    def profile(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        if not self.training:
            return features
        if self._current_iter == -1:
            self._current_iter = 0
            self._last_eviction_iter = self._current_iter
        self._current_iter += 1
        if self._input_history_buffer_size == -1:
            self._init_history_buffers(features)
        for f in features.values():
            v = f.values()
            n = v.numel()
            if self._current_history_buffer_offset + n > self._input_history_buffer_size:
                self._coalesce_history()
            if n > self._input_history_buffer_size:
                v = v[: self._input_history_buffer_size]
                n = v.numel()
            o = self._current_history_buffer_offset
            self._history_ids[o : o + n] = v
            md: Dict[str, torch.Tensor] = {}
            self._eviction_policy.record_history_metadata(self._current_iter, v, md)
            for k, t in md.items():
                self._history_extra[k][o : o + n] = t
            self._current_history_buffer_offset += n
        if self._current_iter - self._last_eviction_iter == self._eviction_interval:
            self._coalesce_history()
            self._last_eviction_iter = self._current_iter
        return features

    def remap(self, features: Dict[str, JaggedTensor]) -> Dict[str, JaggedTensor]:
        out: Dict[str, JaggedTensor] = {}
        for name, f in features.items():
            v = f.values()
            pos = torch.searchsorted(self._mch_sorted_raw_ids, v).clamp(max=self._n_slots - 1)
            hit = self._mch_sorted_raw_ids[pos] == v
            slot = self._mch_remapped_ids_mapping[pos]
            if self._mch_size > 0:
                h = self._mch_hash_func(v, self._mch_size) if self._mch_hash_func is not None else torch.remainder(v, self._mch_size)
                miss = self._n_slots + h
            else:
                miss = torch.remainder(v, self._zch_size)
            remapped = torch.where(hit, slot, miss) + self._output_global_offset
            out[name] = JaggedTensor(values=remapped, lengths=f.lengths(), offsets=f.offsets(), weights=f.weights_or_none())
        return out

    def output_size(self) -> int:
        return self._zch_size

    def input_size(self) -> int:
        return self._input_hash_size

    def buckets(self) -> int:
        return self._buckets

    def open_slots(self) -> torch.Tensor:
        return (self._mch_sorted_raw_ids == torch.iinfo(torch.int64).max).sum().view(1)

    def evict(self) -> Optional[torch.Tensor]:
        if self._evicted:
            self._evicted = False
            return self._evicted_emb_indices
        return None

    def rebuild_with_output_id_range(self, output_id_range: Tuple[int, int], output_segments: List[int], device: Optional[torch.device] = None) -> "MCHManagedCollisionModule":
        new_zch_size = output_id_range[1] - output_id_range[0]
        return type(self)(name=self._name, zch_size=new_zch_size, device=device or self.device, eviction_policy=self._eviction_policy,
                          eviction_interval=self._eviction_interval, input_hash_size=self._input_hash_size, input_hash_func=self._input_hash_func,
                          mch_size=int(self._mch_size * new_zch_size / self._zch_size) if self._mch_size else None, mch_hash_func=self._mch_hash_func,
                          output_global_offset=output_id_range[0], output_segments=output_segments, buckets=self._buckets)


class ManagedCollisionCollection(nn.Module):
    """One ManagedCollisionModule per table; remaps a whole KJT (features map to tables through the embedding configs)."""

    def __init__(self, managed_collision_modules: Dict[str, ManagedCollisionModule], embedding_configs, need_preprocess: bool = True) -> None:
        super().__init__()
        self.need_preprocess = need_preprocess  # False: the ids arrive already pre-processed (e.g. hashed by the data pipeline); the modules skip ``preprocess``
        self._managed_collision_modules = nn.ModuleDict(managed_collision_modules)
        self._embedding_configs = embedding_configs
        self._feature_to_table: Dict[str, str] = {f: c.name for c in embedding_configs for f in c.feature_names}
        self._table_to_features: Dict[str, List[str]] = {c.name: list(c.feature_names) for c in embedding_configs}
        self._table_feature_splits = [len(c.feature_names) for c in embedding_configs if c.name in managed_collision_modules]
        self._compute_need_preprocess = False
        table_to_config = {c.name: c for c in embedding_configs}
        for name, config in table_to_config.items():
            if name not in managed_collision_modules:
                raise ValueError(f"Table {name} is not present in managed_collision_modules")
            assert managed_collision_modules[name].output_size() == config.num_embeddings, (
                f"max_output_id in managed collision module for {name} must match {config.num_embeddings}")
        self._features_order: List[str] = [f for c in embedding_configs for f in c.feature_names]
        for m in self._managed_collision_modules.values():
            m._need_preprocess = need_preprocess

    def embedding_configs(self):
        return self._embedding_configs

    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        jt = features.to_dict()
        out: Dict[str, JaggedTensor] = {}
        for table, mc in self._managed_collision_modules.items():
            feats = {f: jt[f] for f in self._table_to_features[table] if f in jt}
            if feats:
                out.update(mc(feats))
        keys = [k for k in features.keys() if k in out]
        return KeyedJaggedTensor(keys=keys, values=torch.cat([out[k].values() for k in keys]), lengths=torch.cat([out[k].lengths() for k in keys]),
                                 weights=torch.cat([out[k].weights() for k in keys]) if features.weights_or_none() is not None else None, stride=features.stride())

    def evict(self) -> Dict[str, Optional[torch.Tensor]]:
        return {t: mc.evict() for t, mc in self._managed_collision_modules.items()}

    def open_slots(self) -> Dict[str, torch.Tensor]:
        return {t: mc.open_slots() for t, mc in self._managed_collision_modules.items()}
