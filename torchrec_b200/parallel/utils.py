"""Small helpers shared by sharders, DMP and user code (reference torchrec/distributed/utils.py:55-895)."""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Set, Tuple, Type, TypeVar, Union

import torch
from torch import nn

from ..ops.tbe import OptimType
from .logger import ForkedPdb  # noqa: F401  (reference keeps it in utils)
from .types import ParameterSharding, ShardedModule

_T = TypeVar("_T")


def get_device_type() -> str:
    return "cuda" if torch.cuda.is_available() else "cpu"


def get_class_name(obj: object) -> str:
    return f"{type(obj).__module__}.{type(obj).__qualname__}"


def assert_instance(obj: object, t: Type[_T]) -> _T:
    assert isinstance(obj, t), f"Got {get_class_name(obj)}"
    return obj


def none_throws(optional: Optional[_T], message: str = "Unexpected `None`") -> _T:
    if optional is None:
        raise AssertionError(message)
    return optional


def append_prefix(prefix: str, name: str) -> str:
    if prefix != "" and name != "":
        return prefix + "." + name
    return prefix + name


def filter_state_dict(state_dict: "OrderedDict[str, torch.Tensor]", name: str) -> "OrderedDict[str, torch.Tensor]":
    """Entries under ``name.`` with that prefix removed."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    pre = name + "."
    for k, v in state_dict.items():
        if k.startswith(pre):
            out[k[len(pre):]] = v
    return out


def add_prefix_to_state_dict(state_dict: Dict[str, Any], prefix: str) -> None:
    """In place: ``k -> prefix + k`` (also in the ``_metadata`` side table, like torch's own helpers)."""
    for k in sorted(state_dict.keys()):
        state_dict[prefix + k] = state_dict.pop(k)
    meta = getattr(state_dict, "_metadata", None)
    if meta is not None:
        for k in sorted(meta.keys()):
            meta[prefix + k if k else prefix.rstrip(".")] = meta.pop(k)


def get_unsharded_module_names(model: nn.Module) -> List[str]:
    """Top-most module paths that contain no ShardedModule (what DDP / FSDP may wrap)."""
    names: Set[str] = set()

    def helper(m: nn.Module, path: str) -> bool:
        if isinstance(m, ShardedModule):
            return True
        child_sharded = [(n, helper(c, append_prefix(path, n))) for n, c in m.named_children()]
        if any(s for _, s in child_sharded):
            for n, s in child_sharded:
                if not s:
                    names.add(append_prefix(path, n))
            return True
        return False

    if not helper(model, ""):
        names.add("")
    return sorted(names)


class sharded_model_copy:
    """Context manager for ``copy.deepcopy`` of a sharded model onto another device (``device=None`` keeps tensors where they are but
    still skips process groups): ``with sharded_model_copy("cpu"): m2 = copy.deepcopy(m)``. Reference utils.py:188-248."""

    def __init__(self, device: Optional[Union[str, int, torch.device]]) -> None:
        self.device = device

    def __enter__(self) -> None:
        self._tensor_deepcopy = torch.Tensor.__deepcopy__
        self._param_deepcopy = nn.Parameter.__deepcopy__
        dev = self.device

        def tensor_copy(t: torch.Tensor, memo: Dict[int, Any]) -> torch.Tensor:
            if id(t) in memo:
                return memo[id(t)]
            out = t.detach().to(dev) if dev is not None else t.detach().clone()
            if isinstance(t, nn.Parameter):
                out = nn.Parameter(out, requires_grad=t.requires_grad)
            memo[id(t)] = out
            return out

        torch.Tensor.__deepcopy__ = tensor_copy  # type: ignore[assignment]
        nn.Parameter.__deepcopy__ = tensor_copy  # type: ignore[assignment]
        import torch.distributed as dist

        self._pg_deepcopy = getattr(dist.ProcessGroup, "__deepcopy__", None)
        dist.ProcessGroup.__deepcopy__ = lambda pg, memo: pg  # type: ignore[attr-defined]

    def __exit__(self, *exc: Any) -> None:
        import torch.distributed as dist

        torch.Tensor.__deepcopy__ = self._tensor_deepcopy  # type: ignore[assignment]
        nn.Parameter.__deepcopy__ = self._param_deepcopy  # type: ignore[assignment]
        if self._pg_deepcopy is None:
            del dist.ProcessGroup.__deepcopy__  # type: ignore[attr-defined]
        else:
            dist.ProcessGroup.__deepcopy__ = self._pg_deepcopy  # type: ignore[attr-defined]


def copy_to_device(module: nn.Module, current_device: torch.device, to_device: torch.device) -> nn.Module:
    """Deep copy of ``module`` with every tensor on ``current_device`` moved to ``to_device`` (meta -> empty)."""
    with sharded_model_copy(device=None):
        new = copy.deepcopy(module)
    for m in new.modules():
        for name, p in list(m._parameters.items()):
            if p is not None and p.device == current_device:
                data = torch.empty_like(p, device=to_device) if p.is_meta else p.detach().to(to_device)
                m._parameters[name] = nn.Parameter(data, requires_grad=p.requires_grad)
        for name, b in list(m._buffers.items()):
            if b is not None and b.device == current_device:
                m._buffers[name] = torch.empty_like(b, device=to_device) if b.is_meta else b.to(to_device)
    return new


class CopyableMixin(nn.Module):
    """``module.copy(device)``: deep copy onto ``device`` (used by inference replication across local GPUs)."""

    def copy(self, device: torch.device) -> nn.Module:
        cur = next((p.device for p in self.parameters()), next((b.device for b in self.buffers()), torch.device("cpu")))
        return copy_to_device(self, cur, torch.device(device))


_OPT_MAP = {
    "SGD": OptimType.EXACT_SGD, "LarsSGD": OptimType.LARS_SGD, "LAMB": OptimType.LAMB, "PartialRowWiseLAMB": OptimType.PARTIAL_ROWWISE_LAMB,
    "Adam": OptimType.ADAM, "AdamW": OptimType.ADAMW, "PartialRowWiseAdam": OptimType.PARTIAL_ROWWISE_ADAM, "Adagrad": OptimType.EXACT_ADAGRAD,
    "RowWiseAdagrad": OptimType.EXACT_ROWWISE_ADAGRAD, "Lion": OptimType.LION,
}


def optimizer_type_to_emb_opt_type(optimizer_class: Type[torch.optim.Optimizer]) -> Optional[OptimType]:
    """torch / torchrec_b200.optim optimizer class -> fused TBE optimizer (reference utils.py:357-378)."""
    name = optimizer_class.__name__
    if name not in _OPT_MAP:
        raise ValueError(f"Cannot cast {optimizer_class} to a fused embedding optimizer")
    return _OPT_MAP[name]


def emb_opt_type_to_optimizer_class(optim: OptimType) -> Type[torch.optim.Optimizer]:
    from ..optim import optimizers as O
    from ..optim.rowwise_adagrad import RowWiseAdagrad

    table = {OptimType.EXACT_SGD: O.SGD, OptimType.LARS_SGD: O.LarsSGD, OptimType.LAMB: O.LAMB, OptimType.PARTIAL_ROWWISE_LAMB: O.PartialRowWiseLAMB,
             OptimType.ADAM: O.Adam, OptimType.PARTIAL_ROWWISE_ADAM: O.PartialRowWiseAdam, OptimType.EXACT_ADAGRAD: O.Adagrad,
             OptimType.EXACT_ROWWISE_ADAGRAD: RowWiseAdagrad}
    if optim not in table:
        raise ValueError(f"no optimizer class for {optim}")
    return table[optim]


def merge_fused_params(fused_params: Optional[Dict[str, Any]] = None, param_fused_params: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    """Per-parameter (``apply_optimizer_in_backward``) settings override the sharder-level ``fused_params``."""
    merged = dict(fused_params or {})
    merged.update(param_fused_params or {})
    return merged


def add_params_from_parameter_sharding(fused_params: Optional[Dict[str, Any]], parameter_sharding: ParameterSharding) -> Dict[str, Any]:
    """Plan-level per-table knobs (cache params, rounding, bounds checks, kernel output dtype) folded into the fused params."""
    fp = dict(fused_params or {})
    cp = getattr(parameter_sharding, "cache_params", None)
    if cp is not None:
        for src, dst in (("algorithm", "cache_algorithm"), ("load_factor", "cache_load_factor"), ("reserved_memory", "cache_reserved_memory"),
                         ("precision", "cache_precision"), ("prefetch_pipeline", "prefetch_pipeline"), ("multipass_prefetch_config", "multipass_prefetch_config")):
            v = getattr(cp, src, None)
            if v is not None:
                fp[dst] = v
    for attr in ("enforce_hbm", "stochastic_rounding", "bounds_check_mode", "output_dtype"):
        v = getattr(parameter_sharding, attr, None)
        if v is not None:
            fp[attr] = v
    return fp


def init_parameters(module: nn.Module, device: torch.device) -> None:
    """Materialise meta parameters / buffers of ``module`` on ``device`` and run ``reset_parameters`` where available."""
    @torch.no_grad()
    def init(m: nn.Module) -> None:
        had_meta = False
        for name, p in list(m._parameters.items()):
            if p is not None and p.is_meta:
                m._parameters[name] = nn.Parameter(torch.empty_like(p, device=device), requires_grad=p.requires_grad)
                had_meta = True
        for name, b in list(m._buffers.items()):
            if b is not None and b.is_meta:
                m._buffers[name] = torch.zeros_like(b, device=device)
        if had_meta and hasattr(m, "reset_parameters"):
            m.reset_parameters()  # type: ignore[operator]

    module.apply(init)


def weights_bytes_in_emb_kernel(emb: nn.Module) -> int:
    """HBM / host bytes held by the table-batched kernels under ``emb`` (weights only)."""
    total = 0
    for m in emb.modules():
        w = getattr(m, "weights", None)
        if isinstance(w, torch.Tensor) and hasattr(m, "embedding_specs"):
            total += w.numel() * w.element_size()
    return total


def convert_to_fbgemm_types(fused_params: Dict[str, Any]) -> Dict[str, Any]:
    """``cache_precision`` / ``weights_precision`` / ``output_dtype`` given as ``DataType`` become the row-format names this framework's
    kernels key on (the reference converts to FBGEMM's ``SparseType``)."""
    from ..modules.embedding_configs import DataType, data_type_to_sparse_type

    for key in ("cache_precision", "weights_precision", "output_dtype"):
        if isinstance(fused_params.get(key), DataType):
            fused_params[key] = data_type_to_sparse_type(fused_params[key])
    return fused_params


def maybe_annotate_embedding_event(event: Any, module_fqn: Optional[str], sharding_type: Optional[str]):
    """A profiler range ``[<event>]_[<module>]_[<sharding type>]`` around a stage of a sharded embedding module (no-op without both names)."""
    import contextlib

    from torch.autograd.profiler import record_function

    if module_fqn and sharding_type:
        return record_function(f"[{event.value}]_[{module_fqn}]_[{sharding_type}]")
    return contextlib.nullcontext()


def create_global_tensor_shape_stride_from_metadata(parameter_sharding: ParameterSharding, devices_per_node: Optional[int] = None) -> Tuple[torch.Size, Tuple[int, int]]:
    """Global shape and (row-major) stride of a table from the shards of its plan entry; grid shards need the ranks per node."""
    from .types import ShardingType

    shards = parameter_sharding.sharding_spec.shards if parameter_sharding.sharding_spec is not None else []
    st = parameter_sharding.sharding_type
    size = None
    if st in (ShardingType.COLUMN_WISE.value, ShardingType.TABLE_COLUMN_WISE.value):
        size = torch.Size([shards[0].shard_sizes[0], sum(s.shard_sizes[1] for s in shards)])
    elif st in (ShardingType.ROW_WISE.value, ShardingType.TABLE_ROW_WISE.value):
        size = torch.Size([sum(s.shard_sizes[0] for s in shards), shards[0].shard_sizes[1]])
    elif st == ShardingType.TABLE_WISE.value:
        size = torch.Size(shards[0].shard_sizes)
    elif st == ShardingType.GRID_SHARD.value:
        assert devices_per_node is not None, "grid shards: the global shape needs the number of ranks per node"
        size = torch.Size([shards[0].shard_sizes[0] * devices_per_node, shards[0].shard_sizes[1] * (len(shards) // devices_per_node)])
    if size is None:
        return torch.Size([0, 0]), (0, 1)
    return size, (size[1], 1)


def get_bucket_metadata_from_shard_metadata(shards: List[Any], num_buckets: int):
    """A row-wise sharded table cut into ``num_buckets`` equal buckets (ZCH: buckets never straddle shards): buckets per shard, index of
    every shard's first bucket, rows per bucket."""
    from .types import ShardingBucketMetadata

    assert len(shards) > 0, "Shards cannot be empty"
    table_size = shards[-1].shard_offsets[0] + shards[-1].shard_sizes[0]
    assert table_size % num_buckets == 0, f"Table size '{table_size}' must be divisible by num_buckets '{num_buckets}'"
    bucket_size = table_size // num_buckets
    meta = ShardingBucketMetadata(num_buckets_per_shard=[], bucket_offsets_per_shard=[], bucket_size=bucket_size)
    offset = 0
    for shard in shards:
        assert len(shard.shard_offsets) == 1 or shard.shard_offsets[1] == 0, \
            f"Shard shard_offsets[1] '{shard.shard_offsets[1]}' is not 0. Table should be only row-wise sharded for bucketization"
        assert shard.shard_sizes[0] % bucket_size == 0, f"Shard size[0] '{shard.shard_sizes[0]}' is not divisible by bucket size '{bucket_size}'"
        n = shard.shard_sizes[0] // bucket_size
        meta.num_buckets_per_shard.append(n)
        meta.bucket_offsets_per_shard.append(offset)
        offset += n
    return meta


def modify_input_for_feature_processor(features: Any, feature_processors: Any, is_collection: bool) -> None:
    """Run the input-side part of the feature processors BEFORE the input dist, in place on the KJT (row-wise sharding splits bags
    across ranks, so what depends on the position inside a bag must be computed while the bag is whole): the KJT gets a weights tensor
    if it has none, then ``pre_process_input(kjt)`` of the collection, or of every per-feature processor on its feature's slice."""
    with torch.no_grad():
        if features.weights_or_none() is None:
            features._weights = torch.zeros_like(features.values(), dtype=torch.float32)
        if is_collection:
            if hasattr(feature_processors, "pre_process_input"):
                feature_processors.pre_process_input(features)
            return
        for feature in features.keys():
            if feature in feature_processors and hasattr(feature_processors[feature], "pre_process_input"):
                feature_processors[feature].pre_process_input(features[feature])


class EmbeddingQuantizationUtils:
    """Temporarily hold the table-batched kernels of a sharded model in a narrower float type (e.g. fp16 while a memory-hungry
    evaluation runs) and restore them: ``quantize_embedding_modules(model, DataType.FP16)`` converts the weights buffers of every kernel
    (smallest first, so the peak is one extra copy of the largest table group), ``recreate_embedding_modules(model)`` converts them
    back to the precision they had. Optimizer states are not touched."""

    _DTYPES = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}

    def __init__(self) -> None:
        self._original: Dict[nn.Module, torch.dtype] = {}

    @staticmethod
    def _kernels(module: nn.Module) -> List[nn.Module]:
        ks = [m for m in module.modules() if hasattr(m, "embedding_specs") and isinstance(getattr(m, "weights", None), torch.Tensor)]
        return sorted(ks, key=weights_bytes_in_emb_kernel)

    @staticmethod
    def _convert(k: nn.Module, dtype: torch.dtype, use_cpu_turnaround_optimization: bool) -> None:
        w = k.weights
        if w.dtype == dtype:
            return
        with torch.no_grad():
            new = w.detach().cpu().to(dtype).to(w.device) if use_cpu_turnaround_optimization else w.detach().to(dtype)
        if isinstance(w, nn.Parameter):
            k.weights = nn.Parameter(new, requires_grad=w.requires_grad)
        else:
            k._buffers["weights"] = new

    def quantize_embedding_modules(self, module: nn.Module, converted_dtype: Any, use_cpu_turnaround_optimization: bool = False) -> None:
        from ..modules.embedding_configs import data_type_to_sparse_type

        dtype = self._DTYPES[data_type_to_sparse_type(converted_dtype)]
        for k in self._kernels(module):
            self._original.setdefault(k, k.weights.dtype)
            self._convert(k, dtype, use_cpu_turnaround_optimization)

    def recreate_embedding_modules(self, module: nn.Module, use_cpu_turnaround_optimization: bool = False) -> None:
        for k in self._kernels(module):
            if k in self._original:
                self._convert(k, self._original[k], use_cpu_turnaround_optimization)
