"""Synthetic training batches for tests and benchmarks.

Reference: ``torchrec/distributed/test_utils/model_input.py`` - ``ModelInput`` :23-845 (``generate`` :387, ``generate_global_and_local_batches`` :204,
``generate_local_batches`` :331, ``create_standard_kjt`` :735, power-law ids :496), ``VariableBatchModelInput`` :848, ``TdModelInput`` :1062.
A batch = dense features + an unweighted KJT + an optional weighted KJT + labels. Lengths are uniform in ``[0, 2 * pooling)`` (mean = the requested
pooling factor); ids are uniform, or Zipf-like with ``power_law_alpha`` (hot rows - what makes caches and dedup matter).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch

from ...sparse.jagged_tensor import KeyedJaggedTensor
from ...streamable import Pipelineable


def _features_of(tables: Optional[Sequence[Any]]) -> List[Tuple[str, int]]:
    return [(f, t.num_embeddings) for t in (tables or []) for f in t.feature_names]


@dataclass
class ModelInput(Pipelineable):
    float_features: torch.Tensor
    idlist_features: Optional[KeyedJaggedTensor]
    idscore_features: Optional[KeyedJaggedTensor]
    label: torch.Tensor
    dummy: Optional[List[torch.Tensor]] = None

    # ---- Pipelineable ----
    def to(self, device: torch.device, non_blocking: bool = False, data_copy_stream: Optional[torch.Stream] = None) -> "ModelInput":
        mv = lambda t: None if t is None else t.to(device=device, non_blocking=non_blocking)  # noqa: E731
        if data_copy_stream is not None and device.type == "cuda":
            with torch.cuda.stream(data_copy_stream):
                return ModelInput(mv(self.float_features), mv(self.idlist_features), mv(self.idscore_features), mv(self.label),
                                  None if self.dummy is None else [mv(t) for t in self.dummy])
        return ModelInput(mv(self.float_features), mv(self.idlist_features), mv(self.idscore_features), mv(self.label),
                          None if self.dummy is None else [mv(t) for t in self.dummy])

    def record_stream(self, stream: torch.Stream) -> None:
        for t in (self.float_features, self.label, *(self.dummy or [])):
            if t.is_cuda:
                t.record_stream(stream)
        for k in (self.idlist_features, self.idscore_features):
            if k is not None:
                k.record_stream(stream)

    def pin_memory(self) -> "ModelInput":
        pin = lambda t: None if t is None else t.pin_memory()  # noqa: E731
        return ModelInput(pin(self.float_features), pin(self.idlist_features), pin(self.idscore_features), pin(self.label),
                          None if self.dummy is None else [pin(t) for t in self.dummy])

    def size_in_bytes(self) -> int:
        n = self.float_features.numel() * self.float_features.element_size() + self.label.numel() * self.label.element_size()
        for k in (self.idlist_features, self.idscore_features):
            if k is not None:
                for t in (k.values(), k.lengths(), k.weights_or_none()):
                    if t is not None:
                        n += t.numel() * t.element_size()
        return n + sum(t.numel() * t.element_size() for t in (self.dummy or []))

    # ---- generation ----
    @staticmethod
    def _generate_power_law_indices(alpha: float, num_indices: int, num_embeddings: int, dtype: torch.dtype, device: Optional[torch.device],
                                    generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """Inverse-CDF sampling of ``P(rank k) ~ k^-alpha`` over ``[1, num_embeddings]``, then a fixed pseudo-random relabeling so hot ids are spread
        over the table instead of clustered at row 0 (row-wise shards would otherwise see all the traffic on rank 0)."""
        u = torch.rand(num_indices, generator=generator, device=generator.device if generator is not None else None).double().clamp(min=1e-12)
        n = float(num_embeddings)
        if abs(alpha - 1.0) < 1e-6:
            k = torch.exp(u * torch.log(torch.tensor(n + 1.0, dtype=torch.double)))
        else:
            a = 1.0 - alpha
            k = ((u * ((n + 1.0) ** a - 1.0)) + 1.0) ** (1.0 / a)
        k = (k.floor().long() - 1).clamp(0, num_embeddings - 1)
        k = (k * 2654435761 + 12345) % num_embeddings
        return k.to(dtype).to(device) if device is not None else k.to(dtype)

    @staticmethod
    def _create_features_lengths_indices(batch_size: int, features: List[Tuple[str, int]], pooling: List[int], max_lengths: Optional[List[Optional[int]]], use_offsets: bool,
                                         device: Optional[torch.device], indices_dtype: torch.dtype, lengths_dtype: torch.dtype, all_zeros: bool, weighted: bool,
                                         power_law_alpha: Optional[float], generator: Optional[torch.Generator]) -> KeyedJaggedTensor:
        lengths, values, weights = [], [], []
        for i, (name, hash_size) in enumerate(features):
            hi = 2 * pooling[i] + 1
            ln = torch.randint(0, max(hi, 1), (batch_size,), generator=generator).to(lengths_dtype)
            if max_lengths is not None and max_lengths[i] is not None:
                ln = ln.clamp(max=int(max_lengths[i]))
            n = int(ln.sum())
            if all_zeros:
                v = torch.zeros(n, dtype=indices_dtype)
            elif power_law_alpha is not None:
                v = ModelInput._generate_power_law_indices(power_law_alpha, n, hash_size, indices_dtype, None, generator)
            else:
                v = torch.randint(0, hash_size, (n,), generator=generator).to(indices_dtype)
            lengths.append(ln)
            values.append(v)
            if weighted:
                weights.append(torch.rand(n, generator=generator))
        return ModelInput._assemble_kjt([f for f, _ in features], torch.cat(lengths) if lengths else torch.zeros(0, dtype=lengths_dtype),
                                        torch.cat(values) if values else torch.zeros(0, dtype=indices_dtype), torch.cat(weights) if weights else None,
                                        use_offsets, device, lengths_dtype)

    @staticmethod
    def _assemble_kjt(keys: List[str], lengths: torch.Tensor, values: torch.Tensor, weights: Optional[torch.Tensor], use_offsets: bool,
                      device: Optional[torch.device], offsets_dtype: torch.dtype = torch.int64) -> KeyedJaggedTensor:
        mv = (lambda t: t if t is None or device is None else t.to(device))
        if use_offsets:
            offsets = torch.cat([torch.zeros(1, dtype=offsets_dtype), torch.cumsum(lengths, 0).to(offsets_dtype)])
            return KeyedJaggedTensor(keys=keys, values=mv(values), offsets=mv(offsets), weights=mv(weights))
        return KeyedJaggedTensor(keys=keys, values=mv(values), lengths=mv(lengths), weights=mv(weights))

    @staticmethod
    def create_standard_kjt(batch_size: int, tables: Sequence[Any], pooling_avg: int = 10, tables_pooling: Optional[List[int]] = None, weighted: bool = False,
                            max_feature_lengths: Optional[List[int]] = None, use_offsets: bool = False, device: Optional[torch.device] = None,
                            indices_dtype: torch.dtype = torch.int64, offsets_dtype: torch.dtype = torch.int64, lengths_dtype: torch.dtype = torch.int64,
                            all_zeros: bool = False, power_law_alpha: Optional[float] = None, generator: Optional[torch.Generator] = None) -> KeyedJaggedTensor:
        feats = _features_of(tables)
        pooling: List[int] = []
        for ti, t in enumerate(tables):
            pooling += [tables_pooling[ti] if tables_pooling is not None else pooling_avg] * len(t.feature_names)
        ml: Optional[List[Optional[int]]] = None
        if max_feature_lengths is not None:
            ml = [max_feature_lengths[i] if i < len(max_feature_lengths) else None for i in range(len(feats))]
        return ModelInput._create_features_lengths_indices(batch_size, feats, pooling, ml, use_offsets, device, indices_dtype, lengths_dtype, all_zeros, weighted,
                                                           power_law_alpha, generator)

    @classmethod
    def generate(cls, batch_size: int = 1, tables: Optional[Sequence[Any]] = None, weighted_tables: Optional[Sequence[Any]] = None, num_float_features: int = 16,
                 pooling_avg: int = 10, tables_pooling: Optional[List[int]] = None, max_feature_lengths: Optional[List[int]] = None, use_offsets: bool = False,
                 device: Optional[torch.device] = None, indices_dtype: torch.dtype = torch.int64, offsets_dtype: torch.dtype = torch.int64,
                 lengths_dtype: torch.dtype = torch.int64, all_zeros: bool = False, pin_memory: bool = False, power_law_alpha: Optional[float] = None,
                 num_dummy_tensor: int = 0, generator: Optional[torch.Generator] = None) -> "ModelInput":
        gen_dev = None if pin_memory else device
        kw = dict(pooling_avg=pooling_avg, use_offsets=use_offsets, device=gen_dev, indices_dtype=indices_dtype, offsets_dtype=offsets_dtype, lengths_dtype=lengths_dtype,
                  all_zeros=all_zeros, power_law_alpha=power_law_alpha, generator=generator)
        idlist = cls.create_standard_kjt(batch_size, tables, tables_pooling=tables_pooling, max_feature_lengths=max_feature_lengths, weighted=False, **kw) if tables else None
        idscore = cls.create_standard_kjt(batch_size, weighted_tables, weighted=True, **kw) if weighted_tables else None
        ff = torch.rand(batch_size, num_float_features, generator=generator)
        label = torch.rand(batch_size, generator=generator)
        dummy = [torch.rand(batch_size, generator=generator) for _ in range(num_dummy_tensor)] or None
        out = cls(ff if gen_dev is None else ff.to(gen_dev), idlist, idscore, label if gen_dev is None else label.to(gen_dev),
                  None if dummy is None else [d if gen_dev is None else d.to(gen_dev) for d in dummy])
        return out.pin_memory() if pin_memory and torch.cuda.is_available() else out

    @classmethod
    def generate_local_batches(cls, world_size: int, batch_size: int = 1, **kwargs: Any) -> List["ModelInput"]:
        return [cls.generate(batch_size=batch_size, **kwargs) for _ in range(world_size)]

    @classmethod
    def generate_global_and_local_batches(cls, world_size: int, batch_size: int = 1, **kwargs: Any) -> Tuple["ModelInput", List["ModelInput"]]:
        """Local batches + their concatenation (what an unsharded golden model sees: the global batch in rank order)."""
        locals_ = cls.generate_local_batches(world_size, batch_size, **kwargs)
        return cls.concat(locals_), locals_

    @classmethod
    def concat(cls, batches: List["ModelInput"]) -> "ModelInput":
        def cat_kjt(kjts: List[Optional[KeyedJaggedTensor]]) -> Optional[KeyedJaggedTensor]:
            if kjts[0] is None:
                return None
            keys = kjts[0].keys()
            F = len(keys)
            vals, lens, ws = [], [], []
            for f in range(F):
                for k in kjts:
                    B = k.stride()
                    off = k.offsets()
                    lo, hi = int(off[f * B]), int(off[(f + 1) * B])
                    vals.append(k.values()[lo:hi])
                    lens.append(k.lengths()[f * B : (f + 1) * B])
                    if k.weights_or_none() is not None:
                        ws.append(k.weights()[lo:hi])
            return KeyedJaggedTensor(keys=keys, values=torch.cat(vals), lengths=torch.cat(lens), weights=torch.cat(ws) if ws else None)

        return cls(torch.cat([b.float_features for b in batches]), cat_kjt([b.idlist_features for b in batches]), cat_kjt([b.idscore_features for b in batches]),
                   torch.cat([b.label for b in batches]))


@dataclass
class VariableBatchModelInput(ModelInput):
    """Every feature has its own batch size (``stride_per_key_per_rank``) - the VBE input format."""

    @classmethod
    def generate(cls, batch_size: int = 1, tables: Optional[Sequence[Any]] = None, weighted_tables: Optional[Sequence[Any]] = None, num_float_features: int = 16,
                 pooling_avg: int = 10, world_size: int = 1, device: Optional[torch.device] = None, generator: Optional[torch.Generator] = None, **_: Any) -> "VariableBatchModelInput":  # type: ignore[override]
        def vb_kjt(tbls: Optional[Sequence[Any]], weighted: bool) -> Optional[KeyedJaggedTensor]:
            if not tbls:
                return None
            feats = _features_of(tbls)
            strides = [max(1, int(torch.randint(1, batch_size + 1, (1,), generator=generator))) for _ in feats]
            lengths, values, weights = [], [], []
            for (name, h), b in zip(feats, strides):
                ln = torch.randint(0, 2 * pooling_avg + 1, (b,), generator=generator)
                n = int(ln.sum())
                lengths.append(ln)
                values.append(torch.randint(0, h, (n,), generator=generator))
                if weighted:
                    weights.append(torch.rand(n, generator=generator))
            mv = (lambda t: t if device is None else t.to(device))
            return KeyedJaggedTensor(keys=[f for f, _ in feats], values=mv(torch.cat(values)), lengths=mv(torch.cat(lengths)), weights=mv(torch.cat(weights)) if weights else None,
                                     stride_per_key_per_rank=[[b] for b in strides])

        ff = torch.rand(batch_size, num_float_features, generator=generator)
        label = torch.rand(batch_size, generator=generator)
        return cls(ff if device is None else ff.to(device), vb_kjt(tables, False), vb_kjt(weighted_tables, True), label if device is None else label.to(device))


@dataclass
class TdModelInput(ModelInput):
    """Sparse features as a dict of JaggedTensors (the TensorDict-style input of EBC)."""

    # ``idlist_features`` holds Dict[str, JaggedTensor] here

    @classmethod
    def generate(cls, *args: Any, **kwargs: Any) -> "TdModelInput":  # type: ignore[override]
        base = ModelInput.generate(*args, **kwargs)
        jt = base.idlist_features.to_dict() if base.idlist_features is not None else None
        return cls(base.float_features, jt, base.idscore_features, base.label)
