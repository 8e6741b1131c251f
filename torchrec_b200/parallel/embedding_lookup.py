"""Grouped lookups: every ``GroupedEmbeddingConfig`` of a rank -> one kernel; a KJT in rank-local feature order -> one tensor.

Reference: ``torchrec/distributed/embedding_lookup.py`` - ``GroupedEmbeddingsLookup`` (sequence) :197-508, ``GroupedEmbeddingsUpdate`` :510-534,
``CommOpGradientScaling`` :537-558, ``GroupedPooledEmbeddingsLookup`` :561-1183, ``MetaInferGrouped*`` :1185-1497, ``InferGrouped*`` :1500-1705.

The sharded modules of this framework run ``ShardedLookupEngine`` (``parallel/engine.py``) which fuses lookup and output dist over NVLink; the classes
here are the composable form of the same stage (dist modules of ``dist_data.py`` around them, see ``parallel/sharding/*``), kept for users who build
their own ``EmbeddingSharding``s and for the inference path, where one process drives the lookups of several devices.
"""
from __future__ import annotations

from abc import ABC
from collections import OrderedDict
from typing import Any, Dict, Iterator, List, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch import nn

from ..modules.embedding_configs import DataType, data_type_to_dtype
from ..optim.fused import FusedOptimizerModule
from ..optim.keyed import CombinedOptimizer, KeyedOptimizer
from ..sparse.jagged_tensor import KeyedJaggedTensor
from .batched_embedding_kernel import (
    BatchedDenseEmbedding,
    BatchedDenseEmbeddingBag,
    BatchedFusedEmbedding,
    BatchedFusedEmbeddingBag,
    KeyValueEmbedding,
    KeyValueEmbeddingBag,
    ShardedBatchedFusedEmbedding,
    ShardedBatchedFusedEmbeddingBag,
    ZeroCollisionEmbeddingCache,
    ZeroCollisionKeyValueEmbedding,
    ZeroCollisionKeyValueEmbeddingBag,
)
from .embedding_kernel import BaseEmbedding
from .embedding_types import BaseEmbeddingLookup, EmbeddingComputeKernel, GroupedEmbeddingConfig, InputDistOutputs, KJTList
from .quant_embedding_kernel import QuantBatchedEmbedding, QuantBatchedEmbeddingBag
from .types import LazyAwaitable, ShardingEnv, ShardingType


def fx_wrap_tensor_view2d(x: torch.Tensor, dim0: int, dim1: int) -> torch.Tensor:
    return x.view(dim0, dim1)


def dummy_tensor(sparse_features: KeyedJaggedTensor, dtype: torch.dtype) -> torch.Tensor:
    return torch.empty([0], dtype=dtype, device=sparse_features.device()).view(sparse_features.stride(), 0)


def embeddings_cat_empty_rank_handle(embeddings: List[torch.Tensor], dummy_embs_tensor: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """A rank that owns no table of a sharding still takes part in its collectives: it contributes a 0-width tensor that carries a grad_fn."""
    if not embeddings:
        return dummy_embs_tensor
    return embeddings[0] if len(embeddings) == 1 else torch.cat(embeddings, dim=dim)


def embeddings_cat_empty_rank_handle_inference(embeddings: List[torch.Tensor], dim: int = 0, device: Optional[str] = None,
                                               dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    if not embeddings:
        return torch.empty([0], dtype=dtype, device=torch.device(device) if device is not None else None)
    return embeddings[0] if len(embeddings) == 1 else torch.cat(embeddings, dim=dim)


def _load_state_dict(emb_modules: "nn.ModuleList", state_dict: Dict[str, Any]) -> Tuple[List[str], List[str]]:
    """Copy ``{table}.weight`` entries (tensors or ShardedTensors) into the kernels' table views; returns (missing, unexpected)."""
    unexpected = list(state_dict.keys())
    missing: List[str] = []
    for m in emb_modules:
        for (key, dst), t in zip(m.named_split_embedding_weights(), m.config.embedding_tables):
            if key not in state_dict:
                missing.append(key)
                continue
            if key in unexpected:
                unexpected.remove(key)
            src = state_dict[key]
            if hasattr(src, "local_shards"):
                shards = src.local_shards()
                md = t.local_metadata
                src = next((s.tensor for s in shards if md is None or list(s.metadata.shard_offsets) == list(md.shard_offsets)), shards[0].tensor)
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{key}: expected {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.detach().copy_(src.to(dst.dtype))
        inner = getattr(m, "emb_module", None)
        if inner is not None and hasattr(inner, "load_rows_changed"):
            inner.load_rows_changed()
    return missing, unexpected


class CommOpGradientScaling(torch.autograd.Function):
    """Identity whose backward multiplies by ``scale_gradient_factor``: undoes the gradient division of the output collective for features whose
    rows must not be averaged (weighted feature processors) - reference :537-558."""

    @staticmethod
    def forward(ctx, input_tensor: torch.Tensor, scale_gradient_factor: int) -> torch.Tensor:  # type: ignore[override]
        ctx.scale_gradient_factor = scale_gradient_factor
        return input_tensor

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor) -> Tuple[torch.Tensor, None]:  # type: ignore[override]
        return grad_output.mul(ctx.scale_gradient_factor), None


def _is_fully_sharded(env: Optional[ShardingEnv]) -> bool:
    strat = getattr(env, "sharding_strategy", None)
    return strat is not None and getattr(strat, "name", str(strat)) == "FULLY_SHARDED"


class _GroupedBase(BaseEmbeddingLookup[KeyedJaggedTensor, torch.Tensor]):
    _POOLED = True

    def __init__(self, grouped_configs: List[GroupedEmbeddingConfig], pg: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None,
                 feature_processor: Optional[nn.Module] = None, scale_weight_gradients: bool = True, sharding_type: Optional[ShardingType] = None,
                 env: Optional[ShardingEnv] = None) -> None:
        super().__init__()
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        self._pg = pg
        self._env = env
        self._sharding_type = sharding_type
        self.grouped_configs = grouped_configs
        self._feature_processor = feature_processor
        self._world_size = dist.get_world_size(pg) if pg is not None and dist.is_initialized() else 1
        self._scale_gradient_factor = self._world_size if scale_weight_gradients and _gradient_division_on() else 1
        self._emb_modules: nn.ModuleList = nn.ModuleList([self._create_embedding_kernel(c, self._device, pg, sharding_type, env) for c in grouped_configs])
        self._feature_splits: List[int] = [c.num_features() for c in grouped_configs]
        self._need_prefetch = any(hasattr(getattr(m, "emb_module", None), "prefetch") for m in self._emb_modules)
        out_dtype = torch.float32
        for c in grouped_configs:
            od = (c.fused_params or {}).get("output_dtype")
            if od is not None:
                out_dtype = data_type_to_dtype(od) if isinstance(od, DataType) else od
        self.register_buffer("_dummy_embs_tensor", torch.empty([0], dtype=out_dtype, device=self._device if self._device.type != "meta" else "cpu",
                                                                requires_grad=True), persistent=False)

    # -- kernel choice: compute kernel of the group (reference :233-323 / :642-712) --
    def _create_embedding_kernel(self, config: GroupedEmbeddingConfig, device: torch.device, pg: Optional[dist.ProcessGroup],
                                 sharding_type: Optional[ShardingType], env: Optional[ShardingEnv]) -> BaseEmbedding:
        k = config.compute_kernel
        bag = self._POOLED
        if k == EmbeddingComputeKernel.DENSE:
            return (BatchedDenseEmbeddingBag if bag else BatchedDenseEmbedding)(config, pg, device, sharding_type, env)
        if k in (EmbeddingComputeKernel.FUSED, EmbeddingComputeKernel.FUSED_UVM, EmbeddingComputeKernel.FUSED_UVM_CACHING):
            if _is_fully_sharded(env) and k == EmbeddingComputeKernel.FUSED:
                return (ShardedBatchedFusedEmbeddingBag if bag else ShardedBatchedFusedEmbedding)(config, pg, device, sharding_type, env)
            return (BatchedFusedEmbeddingBag if bag else BatchedFusedEmbedding)(config, pg, device, sharding_type, env)
        if k == EmbeddingComputeKernel.KEY_VALUE:
            return (KeyValueEmbeddingBag if bag else KeyValueEmbedding)(config, pg, device, sharding_type, env)
        if k in (EmbeddingComputeKernel.SSD_VIRTUAL_TABLE, EmbeddingComputeKernel.DRAM_VIRTUAL_TABLE):
            if not bag and k == EmbeddingComputeKernel.DRAM_VIRTUAL_TABLE and any(getattr(t, "enable_embedding_update", False) for t in config.embedding_tables):
                return ZeroCollisionEmbeddingCache(config, pg, device, sharding_type, env, backend_type=k.value)
            return (ZeroCollisionKeyValueEmbeddingBag if bag else ZeroCollisionKeyValueEmbedding)(config, pg, device, sharding_type, env, backend_type=k.value)
        raise ValueError(f"Compute kernel not supported {k}")

    # -- forward --
    def _split(self, sparse_features: KeyedJaggedTensor) -> List[KeyedJaggedTensor]:
        if len(self.grouped_configs) == 0:
            return []
        assert sparse_features is not None
        return sparse_features.split(self._feature_splits)

    def prefetch(self, sparse_features: KeyedJaggedTensor, forward_stream: Optional[torch.cuda.Stream] = None) -> None:
        """Stage the rows of the NEXT batch into the HBM caches of the cached kernels (reference :325-365 / :714-767). Runs on the caller's
        (prefetch) stream; the tensors are marked as used by the forward stream so the allocator does not recycle them early."""
        if not self._need_prefetch:
            return
        for m, f in zip(self._emb_modules, self._split(sparse_features)):
            if hasattr(getattr(m, "emb_module", None), "prefetch"):
                m.prefetch(f, forward_stream)
                if forward_stream is not None and f.values().is_cuda:
                    f.record_stream(forward_stream)

    def state_dict(self, destination: Optional[Dict[str, Any]] = None, prefix: str = "", keep_vars: bool = False) -> Dict[str, Any]:  # type: ignore[override]
        if destination is None:
            destination = OrderedDict()
            destination._metadata = OrderedDict()  # type: ignore[attr-defined]
        for m in self._emb_modules:
            m.state_dict(destination, prefix, keep_vars)
        return destination

    def load_state_dict(self, state_dict: Dict[str, Any], strict: bool = True) -> Any:  # type: ignore[override]
        missing, unexpected = _load_state_dict(self._emb_modules, state_dict)
        if strict and (missing or unexpected):
            raise RuntimeError(f"missing {missing}, unexpected {unexpected}")
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
        for m in self._emb_modules:
            yield from m.named_parameters(prefix, recurse)

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, torch.Tensor]]:
        for m in self._emb_modules:
            yield from m.named_buffers(prefix, recurse)

    def named_parameters_by_table(self) -> Iterator[Tuple[str, nn.Parameter]]:
        for m in self._emb_modules:
            if hasattr(m, "named_parameters_by_table"):
                yield from m.named_parameters_by_table()
            else:
                for name, p in getattr(m, "_param_per_table", {}).items():
                    yield name, p

    def get_named_split_embedding_weights_snapshot(self) -> Iterator[Tuple[str, Any, Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]]:
        for m in self._emb_modules:
            if hasattr(m, "get_named_split_embedding_weights_snapshot"):
                yield from m.get_named_split_embedding_weights_snapshot()

    def fused_optimizers(self) -> List[KeyedOptimizer]:
        return [m.fused_optimizer for m in self._emb_modules if isinstance(m, FusedOptimizerModule)]

    def flush(self) -> None:
        for m in self._emb_modules:
            m.flush()

    def purge(self) -> None:
        for m in self._emb_modules:
            m.purge()

    def get_resize_awaitables(self) -> List[LazyAwaitable[torch.Tensor]]:
        return []

    def register_optim_state_tracker_fn(self, record_fn: Any) -> None:
        for m in self._emb_modules:
            m.init_raw_id_tracker(record_fn, None)


def _gradient_division_on() -> bool:
    from .comm_ops import get_gradient_division

    return bool(get_gradient_division())


class GroupedEmbeddingsLookup(_GroupedBase):
    """Sequence lookup: output ``[sum of lengths of all local features, D]`` - the groups' rows concatenated in feature order (all tables of one
    sequence sharding share ``D`` per group; groups are concatenated along dim 0 when dims match, else dim 1 of the padded view is the caller's job)."""

    _POOLED = False

    def forward(self, sparse_features: KeyedJaggedTensor) -> torch.Tensor:
        embs = [m(f) for m, f in zip(self._emb_modules, self._split(sparse_features))]
        return embeddings_cat_empty_rank_handle(embs, self._dummy_embs_tensor.view(0, 0) if not embs else self._dummy_embs_tensor, dim=0)


class GroupedEmbeddingsUpdate(nn.Module):
    """Direct row writes (``enable_embedding_update`` tables): ``forward(ids KJT whose weights carry the new rows)`` - reference :510-534."""

    def __init__(self, grouped_emb_lookup: GroupedEmbeddingsLookup, pg: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._lookup = grouped_emb_lookup

    @torch.no_grad()
    def forward(self, embeddings: KeyedJaggedTensor) -> None:
        parts = embeddings.split(self._lookup._feature_splits) if self._lookup._feature_splits else []
        for m, f in zip(self._lookup._emb_modules, parts):
            if f.values().numel() == 0:
                continue
            tables = m.split_embedding_weights()
            ftm = m._feature_table_map
            lens = f.lengths().view(len(ftm), -1).sum(1).tolist()
            rows = f.weights().view(f.values().numel(), -1)
            o = 0
            for fi, n in enumerate(lens):
                if n:
                    tables[ftm[fi]].index_copy_(0, f.values()[o : o + n].long(), rows[o : o + n].to(tables[ftm[fi]].dtype))
                o += n
            inner = getattr(m, "emb_module", None)
            if inner is not None and hasattr(inner, "load_rows_changed"):
                inner.load_rows_changed()


class GroupedPooledEmbeddingsLookup(_GroupedBase):
    """Pooled lookup: ``[B, sum of local feature dims]`` - group outputs concatenated along dim 1 (reference :561).

    ``feature_processor`` (a ``BaseGroupedFeatureProcessor``) is applied to the groups with ``has_feature_processor``; their outputs get
    ``CommOpGradientScaling`` so the position-weight gradients are not divided by the world size."""

    _POOLED = True

    def _apply_feature_processor_and_gradient_scaling(self, config: GroupedEmbeddingConfig, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        if not (config.has_feature_processor and self._feature_processor is not None):
            return features
        features = self._feature_processor(features)
        if self._scale_gradient_factor != 1 and features.weights_or_none() is not None and features.weights().requires_grad:
            features._weights = CommOpGradientScaling.apply(features.weights(), self._scale_gradient_factor)
        return features

    def forward(self, sparse_features: KeyedJaggedTensor) -> torch.Tensor:
        embs: List[torch.Tensor] = []
        for c, m, f in zip(self.grouped_configs, self._emb_modules, self._split(sparse_features)):
            embs.append(m(self._apply_feature_processor_and_gradient_scaling(c, f)))
        if not embs:
            return self._dummy_embs_tensor.view(sparse_features.stride() if sparse_features is not None else 0, 0)
        return embeddings_cat_empty_rank_handle(embs, self._dummy_embs_tensor, dim=1)


# ---- inference: one process, several devices -------------------------------------------------------------------------
class _MetaInferBase(BaseEmbeddingLookup[KeyedJaggedTensor, torch.Tensor]):
    """The lookups of ONE device during inference: quantized kernels, no autograd (reference ``MetaInferGrouped*`` :1185 / :1332)."""

    _KERNEL = QuantBatchedEmbeddingBag
    _DIM = 1

    def __init__(self, grouped_configs: List[GroupedEmbeddingConfig], device: Optional[torch.device] = None, feature_processor: Optional[nn.Module] = None,
                 fused_params: Optional[Dict[str, Any]] = None, shard_index: Optional[int] = None) -> None:
        super().__init__()
        self.grouped_configs = grouped_configs
        self._feature_processor = feature_processor
        self._emb_modules: nn.ModuleList = nn.ModuleList([self._KERNEL(c, device=device, fused_params=fused_params, shard_index=shard_index) for c in grouped_configs])
        self._feature_splits = [c.num_features() for c in grouped_configs]
        self.device = device
        self.output_dtype = ((fused_params or {}).get("output_dtype") or torch.float32)

    def get_tbes_to_register(self) -> Dict[nn.Module, GroupedEmbeddingConfig]:
        out: Dict[nn.Module, GroupedEmbeddingConfig] = {}
        for m in self._emb_modules:
            out.update(m.get_tbes_to_register())
        return out

    def forward(self, sparse_features: KeyedJaggedTensor) -> torch.Tensor:
        if not self.grouped_configs:
            return embeddings_cat_empty_rank_handle_inference([], self._DIM, str(self.device) if self.device is not None else None, torch.float32)
        parts = sparse_features.split(self._feature_splits)
        embs = []
        for c, m, f in zip(self.grouped_configs, self._emb_modules, parts):
            if c.has_feature_processor and self._feature_processor is not None:
                f = self._feature_processor(f)
            embs.append(m(f))
        return embeddings_cat_empty_rank_handle_inference(embs, self._DIM)

    def state_dict(self, destination: Optional[Dict[str, Any]] = None, prefix: str = "", keep_vars: bool = False) -> Dict[str, Any]:  # type: ignore[override]
        if destination is None:
            destination = OrderedDict()
            destination._metadata = OrderedDict()  # type: ignore[attr-defined]
        for m in self._emb_modules:
            m.state_dict(destination, prefix, keep_vars)
        return destination

    def load_state_dict(self, state_dict: Dict[str, Any], strict: bool = True) -> Any:  # type: ignore[override]
        missing, unexpected = _load_state_dict(self._emb_modules, state_dict)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
        yield from ()

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, torch.Tensor]]:
        for m in self._emb_modules:
            yield from m.named_buffers(prefix, recurse)

    def flush(self) -> None:
        pass

    def purge(self) -> None:
        pass


class MetaInferGroupedEmbeddingsLookup(_MetaInferBase):
    _KERNEL = QuantBatchedEmbedding
    _DIM = 0


class MetaInferGroupedPooledEmbeddingsLookup(_MetaInferBase):
    _KERNEL = QuantBatchedEmbeddingBag
    _DIM = 1


class InferGroupedLookupMixin(ABC):
    """``forward(InputDistOutputs) -> [tensor per device]``: runs device ``i``'s lookup on features ``i`` (reference :1500-1561)."""

    _embedding_lookups_per_rank: List[_MetaInferBase]

    def forward(self, input_dist_outputs: Union[InputDistOutputs, KJTList, List[KeyedJaggedTensor]]) -> List[torch.Tensor]:
        feats = input_dist_outputs.features if isinstance(input_dist_outputs, InputDistOutputs) else input_dist_outputs
        return [lookup(feats[i]) for i, lookup in enumerate(self._embedding_lookups_per_rank)]

    def state_dict(self, destination: Optional[Dict[str, Any]] = None, prefix: str = "", keep_vars: bool = False) -> Dict[str, Any]:
        if destination is None:
            destination = OrderedDict()
            destination._metadata = OrderedDict()  # type: ignore[attr-defined]
        for lookup in self._embedding_lookups_per_rank:
            lookup.state_dict(destination, prefix, keep_vars)
        return destination

    def load_state_dict(self, state_dict: Dict[str, Any], strict: bool = True) -> Any:
        missing: List[str] = []
        for lookup in self._embedding_lookups_per_rank:
            missing.extend(lookup.load_state_dict(state_dict, strict=False).missing_keys)
        return torch.nn.modules.module._IncompatibleKeys(missing, [])

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
        yield from ()

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, torch.Tensor]]:
        for i, lookup in enumerate(self._embedding_lookups_per_rank):
            yield from lookup.named_buffers(prefix, recurse)

    def get_tbes_to_register(self) -> Dict[nn.Module, GroupedEmbeddingConfig]:
        out: Dict[nn.Module, GroupedEmbeddingConfig] = {}
        for lookup in self._embedding_lookups_per_rank:
            out.update(lookup.get_tbes_to_register())
        return out


def _infer_device(device_type: str, rank: int) -> torch.device:
    return torch.device(device_type, rank) if device_type == "cuda" else torch.device(device_type)


class InferGroupedPooledEmbeddingsLookup(InferGroupedLookupMixin, BaseEmbeddingLookup[InputDistOutputs, List[torch.Tensor]]):
    def __init__(self, grouped_configs_per_rank: List[List[GroupedEmbeddingConfig]], world_size: int, fused_params: Optional[Dict[str, Any]] = None,
                 device: Optional[torch.device] = None, feature_processor: Optional[nn.Module] = None, device_type_from_sharding_infos: Optional[Union[str, Tuple[str, ...]]] = None) -> None:
        nn.Module.__init__(self)
        dt = device_type_from_sharding_infos if isinstance(device_type_from_sharding_infos, str) else (device.type if device is not None else "cpu")
        self._embedding_lookups_per_rank = nn.ModuleList([  # type: ignore[assignment]
            MetaInferGroupedPooledEmbeddingsLookup(grouped_configs_per_rank[r], _infer_device(dt, r) if dt != "meta" else device, feature_processor, fused_params, shard_index=r)
            for r in range(world_size)])


class InferGroupedEmbeddingsLookup(InferGroupedLookupMixin, BaseEmbeddingLookup[InputDistOutputs, List[torch.Tensor]]):
    def __init__(self, grouped_configs_per_rank: List[List[GroupedEmbeddingConfig]], world_size: int, fused_params: Optional[Dict[str, Any]] = None,
                 device: Optional[torch.device] = None, device_type_from_sharding_infos: Optional[Union[str, Tuple[str, ...]]] = None) -> None:
        nn.Module.__init__(self)
        dt = device_type_from_sharding_infos if isinstance(device_type_from_sharding_infos, str) else (device.type if device is not None else "cpu")
        self._embedding_lookups_per_rank = nn.ModuleList([  # type: ignore[assignment]
            MetaInferGroupedEmbeddingsLookup(grouped_configs_per_rank[r], _infer_device(dt, r) if dt != "meta" else device, None, fused_params, shard_index=r)
            for r in range(world_size)])
