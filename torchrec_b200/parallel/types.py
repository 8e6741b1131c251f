"""Core protocols of the model-parallel runtime (reference torchrec/distributed/types.py).

ShardingType / ParameterSharding / ShardingPlan describe *where* table shards live;
Awaitable / LazyAwaitable let collectives overlap with dense compute; ShardedModule is the
3-phase contract (input_dist -> compute -> output_dist) that DistributedModelParallel and the
train pipelines drive; ShardingEnv wraps the process group(s).
"""
from __future__ import annotations

import abc
import operator
from dataclasses import dataclass, field
from enum import Enum, unique
from typing import Any, Callable, Dict, Generic, Iterator, List, Optional, Tuple, Type, TypeVar, Union

import torch
import torch.distributed as dist
from torch import nn
from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardedTensorMetadata  # noqa: F401
from torch.distributed._shard.sharding_spec import EnumerableShardingSpec, ShardingSpec, ShardMetadata  # noqa: F401

from ..streamable import Multistreamable
from ..types import DataType


@unique
class ShardingType(Enum):
    """How one embedding table is partitioned over ranks (reference types.py:125-145)."""

    DATA_PARALLEL = "data_parallel"  # replicated, dense all-reduce
    TABLE_WISE = "table_wise"  # whole table on one rank
    COLUMN_WISE = "column_wise"  # embedding dim split over ranks
    ROW_WISE = "row_wise"  # rows split over all ranks
    TABLE_ROW_WISE = "table_row_wise"  # rows split over the ranks of one node
    TABLE_COLUMN_WISE = "table_column_wise"  # columns split over the ranks of one node
    GRID_SHARD = "grid_shard"  # column shards, each row-split over a node


class PipelineType(Enum):
    NONE = "none"
    TRAIN_BASE = "train_base"
    TRAIN_SPARSE_DIST = "train_sparse_dist"
    TRAIN_PREFETCH_SPARSE_DIST = "train_prefetch_sparse_dist"


class ModuleShardingPlan:
    pass


class CacheAlgorithm(Enum):
    LRU = "lru"
    LFU = "lfu"


class MultiPassPrefetchConfig:
    def __init__(self, num_passes: int = 12, min_splitable_pass_size: int = 256 * 1024 * 1024) -> None:
        self.num_passes = num_passes
        self.min_splitable_pass_size = min_splitable_pass_size


@dataclass
class CacheStatistics(abc.ABC):
    @property
    @abc.abstractmethod
    def expected_lookups(self) -> float:
        ...

    @abc.abstractmethod
    def expected_miss_rate(self, clf: float) -> float:
        ...

    @property
    @abc.abstractmethod
    def cacheability(self) -> float:
        ...


@dataclass
class CacheParams:
    """Host-offloaded (UVM-style) table cache parameters."""

    algorithm: Optional[CacheAlgorithm] = None
    load_factor: Optional[float] = None
    reserved_memory: Optional[float] = None
    precision: Optional[DataType] = None
    prefetch_pipeline: Optional[bool] = None
    stats: Optional[CacheStatistics] = None
    multipass_prefetch_config: Optional[MultiPassPrefetchConfig] = None

    def __hash__(self) -> int:
        return hash((self.algorithm, self.load_factor, self.reserved_memory, self.precision, self.prefetch_pipeline))


@dataclass
class KeyValueParams:
    """Parameters of key-value (SSD / DRAM) backed virtual tables."""

    ssd_storage_directory: Optional[str] = None
    ssd_rocksdb_write_buffer_size: Optional[int] = None
    ssd_rocksdb_shards: Optional[int] = None
    gather_ssd_cache_stats: Optional[bool] = None
    stats_reporter_config: Optional[Any] = None
    use_passed_in_path: bool = True
    l2_cache_size: Optional[int] = None
    max_l1_cache_size: Optional[int] = None
    enable_async_update: Optional[bool] = None
    bulk_init_chunk_size: Optional[int] = None
    lazy_bulk_init_enabled: Optional[bool] = None
    backend_type: Optional[Any] = None
    kv_zch_params: Optional[Any] = None

    def __hash__(self) -> int:
        return hash((self.ssd_storage_directory, self.ssd_rocksdb_write_buffer_size, self.ssd_rocksdb_shards, self.l2_cache_size))


class CommOp(Enum):
    POOLED_EMBEDDINGS_ALL_TO_ALL = "pooled_embeddings_all_to_all"
    POOLED_EMBEDDINGS_REDUCE_SCATTER = "pooled_embeddings_reduce_scatter"
    SEQUENCE_EMBEDDINGS_ALL_TO_ALL = "sequence_embeddings_all_to_all"


class EmbeddingEvent(Enum):
    KJT_SPLITS_DIST = "splits_dist"
    KJT_TENSORS_DIST = "tensors_dist"
    LOOKUP = "lookup"
    OUTPUT_DIST = "output_dist"
    OUTPUT_DIST_WAIT = "output_dist_wait"


QuantizationContext = TypeVar("QuantizationContext")


class NoOpQuantizedCommCodec(Generic[QuantizationContext]):
    def encode(self, input_tensor: torch.Tensor, ctx=None) -> torch.Tensor:
        return input_tensor

    def decode(self, input_grad: torch.Tensor, ctx=None) -> torch.Tensor:
        return input_grad

    def quantized_dtype(self) -> torch.dtype:
        return torch.float

    def calc_quantized_size(self, input_len: int, ctx=None) -> int:
        return input_len

    def create_context(self):
        return None

    def padded_size(self, input_tensor, dim_per_rank, my_rank, qcomm_ctx) -> Tuple[int, int]:
        return input_tensor.shape[0], 0


class QuantizedCommCodec(Generic[QuantizationContext]):
    """Wire-format codec applied around a collective (reference types.py:216-325)."""

    def encode(self, input_tensor: torch.Tensor, ctx: Optional[QuantizationContext] = None) -> torch.Tensor:
        ...

    def decode(self, input_grad: torch.Tensor, ctx: Optional[QuantizationContext] = None) -> torch.Tensor:
        ...

    @property
    def quantized_dtype(self) -> torch.dtype:
        ...

    def calc_quantized_size(self, input_len: int, ctx: Optional[QuantizationContext] = None) -> int:
        ...

    def create_context(self) -> Optional[QuantizationContext]:
        ...


@dataclass
class QuantizedCommCodecs:
    """Forward / backward codecs of one comm op."""

    forward: Any = field(default_factory=NoOpQuantizedCommCodec)
    backward: Any = field(default_factory=NoOpQuantizedCommCodec)


# ---- awaitables -----------------------------------------------------------------------------
W = TypeVar("W")
M = TypeVar("M")
Out = TypeVar("Out")
CompIn = TypeVar("CompIn")
DistOut = TypeVar("DistOut")
ShrdCtx = TypeVar("ShrdCtx", bound=Multistreamable)


class Awaitable(abc.ABC, Generic[W]):
    """Handle on an in-flight result; ``wait()`` blocks (stream-wise) and runs callbacks."""

    def __init__(self) -> None:
        self._callbacks: List[Callable[[W], W]] = []

    @abc.abstractmethod
    def _wait_impl(self) -> W:
        ...

    def wait(self) -> W:
        ret = self._wait_impl()
        for cb in self.callbacks:
            ret = cb(ret)
        return ret

    @property
    def callbacks(self) -> List[Callable[[W], W]]:
        return self._callbacks


class NoWait(Awaitable[W]):
    def __init__(self, obj: W) -> None:
        super().__init__()
        self._obj = obj

    def _wait_impl(self) -> W:
        return self._obj


class _LazyAwaitableMeta(abc.ABCMeta):
    pass


class LazyAwaitable(Awaitable[W], metaclass=_LazyAwaitableMeta):
    """Awaitable that waits itself the first time the result is *used* (any torch function,
    attribute access, indexing, arithmetic) — lets a model call ``ebc(kjt)`` and keep running
    dense compute until the embeddings are really needed (reference types.py:397-590)."""

    def __init__(self) -> None:
        super().__init__()
        self._result: Optional[W] = None

    @staticmethod
    def _wait_async(obj: Any) -> Any:
        if isinstance(obj, LazyAwaitable):
            if obj._result is None:
                obj._result = obj.wait()
            return obj._result
        return obj

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        new_args = torch.fx.node.map_aggregate(args, LazyAwaitable._wait_async)
        new_kwargs = torch.fx.node.map_aggregate(kwargs, LazyAwaitable._wait_async)
        return func(*new_args, **new_kwargs)

    def __getattr__(self, name):
        if name in ("_result", "_callbacks"):
            raise AttributeError(name)
        res = LazyAwaitable._wait_async(self)
        return getattr(res, name)


def _lazy_binop(op):
    def fn(self, other):
        return op(LazyAwaitable._wait_async(self), LazyAwaitable._wait_async(other))

    return fn


def _lazy_rbinop(op):
    def fn(self, other):
        return op(LazyAwaitable._wait_async(other), LazyAwaitable._wait_async(self))

    return fn


for _name, _op in (("add", operator.add), ("sub", operator.sub), ("mul", operator.mul), ("truediv", operator.truediv),
                   ("floordiv", operator.floordiv), ("mod", operator.mod), ("pow", operator.pow), ("matmul", operator.matmul),
                   ("lshift", operator.lshift), ("rshift", operator.rshift), ("and", operator.and_), ("or", operator.or_), ("xor", operator.xor)):
    setattr(LazyAwaitable, f"__{_name}__", _lazy_binop(_op))
    setattr(LazyAwaitable, f"__r{_name}__", _lazy_rbinop(_op))
for _name, _op in (("lt", operator.lt), ("le", operator.le), ("gt", operator.gt), ("ge", operator.ge)):
    setattr(LazyAwaitable, f"__{_name}__", _lazy_binop(_op))
LazyAwaitable.__getitem__ = lambda self, k: LazyAwaitable._wait_async(self)[k]  # type: ignore[assignment]
LazyAwaitable.__len__ = lambda self: len(LazyAwaitable._wait_async(self))  # type: ignore[assignment]
LazyAwaitable.__iter__ = lambda self: iter(LazyAwaitable._wait_async(self))  # type: ignore[assignment]
LazyAwaitable.__neg__ = lambda self: -LazyAwaitable._wait_async(self)  # type: ignore[assignment]
LazyAwaitable.__abs__ = lambda self: abs(LazyAwaitable._wait_async(self))  # type: ignore[assignment]


class LazyNoWait(LazyAwaitable[W]):
    def __init__(self, obj: W) -> None:
        super().__init__()
        self._obj = obj

    def _wait_impl(self) -> W:
        return self._obj


class LazyGetItemMixin(Generic[W]):
    pass


# ---- plan types -------------------------------------------------------------------------------
class ObjectPoolShardingType(Enum):
    ROW_WISE = "row_wise"
    REPLICATED_ROW_WISE = "replicated_row_wise"


@dataclass
class ParameterSharding:
    """Placement of one parameter (table) — reference types.py:753-786.

    ``sharding_spec`` is an ``EnumerableShardingSpec`` listing every shard
    (offsets [row, col], sizes, placement ``rank:r/cuda:r``); ``ranks`` lists owner ranks."""

    sharding_type: str
    compute_kernel: str
    ranks: Optional[List[int]] = None
    sharding_spec: Optional[ShardingSpec] = None
    cache_params: Optional[CacheParams] = None
    enforce_hbm: Optional[bool] = None
    stochastic_rounding: Optional[bool] = None
    bounds_check_mode: Optional[Any] = None
    output_dtype: Optional[DataType] = None
    key_value_params: Optional[KeyValueParams] = None
    use_gpu_rank_broadcast: Optional[bool] = None


class EmbeddingModuleShardingPlan(ModuleShardingPlan, Dict[str, ParameterSharding]):
    """table name -> ParameterSharding."""

    def __str__(self) -> str:
        rows = []
        for name, ps in self.items():
            shards = []
            if ps.sharding_spec is not None:
                for s in ps.sharding_spec.shards:  # type: ignore[attr-defined]
                    shards.append(f"{s.shard_offsets}+{s.shard_sizes}@{s.placement}")
            rows.append(f"  {name}: {ps.sharding_type} {ps.compute_kernel} ranks={ps.ranks} shards=[{', '.join(shards)}]")
        return "\n".join(rows)


@dataclass
class ShardingPlan:
    """module FQN -> module sharding plan (reference types.py:852-900)."""

    plan: Dict[str, ModuleShardingPlan]

    def get_plan_for_module(self, module_path: str) -> Optional[ModuleShardingPlan]:
        return self.plan.get(module_path, None)

    def __str__(self) -> str:
        return "\n\n".join(f"module: {k}\n{v}" for k, v in self.plan.items())


class ObjectPoolShardingPlan(ModuleShardingPlan):
    def __init__(self, sharding_type: ObjectPoolShardingType, inference: bool = False) -> None:
        self.sharding_type = sharding_type
        self.inference = inference


# ---- environments -------------------------------------------------------------------------------
class ShardingEnv:
    """World size / rank / process group abstraction (reference types.py:904-948)."""

    def __init__(self, world_size: int, rank: int, pg: Optional[dist.ProcessGroup] = None, output_dtensor: bool = False) -> None:
        self.world_size = world_size
        self.rank = rank
        self.process_group: Optional[dist.ProcessGroup] = pg
        self.device_mesh = None
        self.output_dtensor = output_dtensor

    @classmethod
    def from_process_group(cls, pg: dist.ProcessGroup, output_dtensor: bool = False) -> "ShardingEnv":
        return cls(dist.get_world_size(pg), dist.get_rank(pg), pg, output_dtensor)

    @classmethod
    def from_local(cls, world_size: int, rank: int) -> "ShardingEnv":
        """Environment without a process group (single-process multi-device inference)."""
        return cls(world_size, rank, None)

    @classmethod
    def from_loopback(cls, world_size: int, rank: int, group) -> "ShardingEnv":
        """Virtual rank ``rank`` of ``world_size`` ranks that all live in THIS process on one device (``sparse_plane.LoopbackGroup``):
        the NVLink plane's kernels run with every "peer" buffer on the local device, the caller drives the ranks in lock step."""
        env = cls(world_size, rank, None)
        env.loopback_group = group
        return env


class ShardingStrategy(Enum):
    DEFAULT = "default"
    PER_MODULE = "per_module"
    FULLY_SHARDED = "fully_sharded"


@dataclass
class DMPCollectionConfig:
    module: Type[nn.Module]
    plan: "ShardingPlan"
    sharding_group_size: int
    node_group_size: Optional[int] = None
    use_inter_host_allreduce: bool = False
    sharding_strategy: ShardingStrategy = ShardingStrategy.PER_MODULE


@dataclass
class DMPCollectionContext(DMPCollectionConfig):
    device_mesh: Any = None
    sharding_pg: Any = None
    replica_pg: Any = None
    modules_to_sync: List[Tuple[nn.Module, nn.Module]] = field(default_factory=list)
    sharded_module: Optional[nn.Module] = None


class ShardingEnv2D(ShardingEnv):
    """2D parallel env: ``sharding_pg`` (model-parallel group) x ``replica_pg`` (reference types.py:1091-1170)."""

    def __init__(self, sharding_pg: dist.ProcessGroup, replica_pg: Optional[dist.ProcessGroup] = None, global_pg: Optional[dist.ProcessGroup] = None, device_mesh=None,
                 node_group_size: Optional[int] = None, use_inter_host_allreduce: bool = False, sharding_strategy: Optional["ShardingStrategy"] = None) -> None:
        # positional order of the reference: (sharding_pg, replica_pg, global_pg, device_mesh, ...)
        assert global_pg is not None, "ShardingEnv2D needs the global process group"
        self.sharding_strategy = sharding_strategy if sharding_strategy is not None else ShardingStrategy.DEFAULT  # how the replica dimension treats the tables (2D: replicate; fully sharded: shard)
        self.world_size = dist.get_world_size(sharding_pg)
        self.global_world_size = dist.get_world_size(global_pg)
        self.rank = dist.get_rank(global_pg)
        self.local_rank = dist.get_rank(sharding_pg)
        self.process_group = global_pg
        self.sharding_pg = sharding_pg
        self.replica_pg = replica_pg
        self.device_mesh = device_mesh
        self.node_group_size = node_group_size
        self.output_dtensor = True
        self.use_inter_host_allreduce = use_inter_host_allreduce
        self.num_sharding_groups = self.global_world_size // self.world_size

    def remap_rank(self, rank: int, sharding_type: ShardingType) -> int:
        """Global rank -> rank inside the sharding group."""
        if self.use_inter_host_allreduce:
            return rank % self.world_size
        return rank // self.num_sharding_groups if sharding_type in (ShardingType.COLUMN_WISE, ShardingType.TABLE_WISE) or True else rank


class NullShardingContext(Multistreamable):
    def record_stream(self, stream: torch.Stream) -> None:
        pass


class NullShardedModuleContext(Multistreamable):
    def record_stream(self, stream: torch.Stream) -> None:
        pass

    def __setattr__(self, key: str, value: Any) -> None:
        raise NotImplementedError()


# ---- module-level contracts ---------------------------------------------------------------------
class FeatureShardingMixIn:
    def feature_names(self) -> List[str]:
        raise NotImplementedError

    def feature_names_per_rank(self) -> List[List[str]]:
        raise NotImplementedError

    def features_per_rank(self) -> List[int]:
        raise NotImplementedError


class ShardedModule(abc.ABC, nn.Module, Generic[CompIn, DistOut, Out, ShrdCtx]):
    """3-phase sharded module: ``input_dist`` -> ``compute`` -> ``output_dist``
    (reference types.py:1184-1340). ``forward`` chains them; train pipelines call the phases on
    different streams."""

    _FORCE_STATE_DICT_LOAD = True

    def __init__(self, qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        super().__init__()
        self._input_dists: List[nn.Module] = []
        self._lookups: List[nn.Module] = []
        self._output_dists: List[nn.Module] = []
        self._qcomm_codecs_registry = qcomm_codecs_registry

    @abc.abstractmethod
    def create_context(self) -> ShrdCtx:
        ...

    @property
    def qcomm_codecs_registry(self) -> Optional[Dict[str, QuantizedCommCodecs]]:
        return self._qcomm_codecs_registry

    @abc.abstractmethod
    def input_dist(self, ctx: ShrdCtx, *input, **kwargs) -> Awaitable[Awaitable[CompIn]]:
        ...

    @abc.abstractmethod
    def compute(self, ctx: ShrdCtx, dist_input: CompIn) -> DistOut:
        ...

    @abc.abstractmethod
    def output_dist(self, ctx: ShrdCtx, output: DistOut) -> LazyAwaitable[Out]:
        ...

    def compute_and_output_dist(self, ctx: ShrdCtx, input: CompIn) -> LazyAwaitable[Out]:
        output = self.compute(ctx, input)
        return self.output_dist(ctx, output)

    def forward(self, *input, **kwargs) -> LazyAwaitable[Out]:
        ctx = self.create_context()
        dist_input = self.input_dist(ctx, *input, **kwargs).wait().wait()
        return self.compute_and_output_dist(ctx, dist_input)

    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        for key, _ in self.named_parameters(prefix):
            yield key

    def extra_repr(self) -> str:
        return ""

    @property
    def unsharded_module_type(self) -> Type[nn.Module]:
        raise NotImplementedError


def get_tensor_size_bytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


class ModuleSharder(abc.ABC, Generic[M]):
    """Knows how to shard one module type (reference types.py:1393-1480)."""

    def __init__(self, qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None) -> None:
        self._qcomm_codecs_registry = qcomm_codecs_registry

    @abc.abstractmethod
    def shard(self, module: M, params: EmbeddingModuleShardingPlan, env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedModule:
        ...

    @property
    @abc.abstractmethod
    def module_type(self) -> Type[M]:
        ...

    @property
    def qcomm_codecs_registry(self) -> Optional[Dict[str, QuantizedCommCodecs]]:
        return self._qcomm_codecs_registry

    def shardable_parameters(self, module: M) -> Dict[str, nn.Parameter]:
        return dict(module.named_parameters())

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [ShardingType.DATA_PARALLEL.value]

    def compute_kernels(self, sharding_type: str, compute_device_type: str) -> List[str]:
        return ["dense"]

    def storage_usage(self, tensor: torch.Tensor, compute_device_type: str, compute_kernel: str) -> Dict[str, int]:
        assert compute_device_type in {"cuda", "cpu", "mtia"}
        storage_map = {"cuda": "HBM", "cpu": "DDR", "mtia": "DDR"}
        return {storage_map[compute_device_type]: get_tensor_size_bytes(tensor)}


class ShardingPlanner(abc.ABC):
    @abc.abstractmethod
    def plan(self, module: nn.Module, sharders: List[ModuleSharder[nn.Module]]) -> ShardingPlan:
        ...

    @abc.abstractmethod
    def collective_plan(self, module: nn.Module, sharders: List[ModuleSharder[nn.Module]], pg: Optional[dist.ProcessGroup] = None) -> ShardingPlan:
        ...


def rank_device(device_type: str, rank: int) -> torch.device:
    if device_type == "cpu":
        return torch.device("cpu")
    return torch.device(f"{device_type}:{rank}")


class ShardingBucketMetadata:
    def __init__(self, num_buckets_per_shard: List[int], bucket_offsets_per_shard: List[int], bucket_size: int) -> None:
        self.num_buckets_per_shard = num_buckets_per_shard
        self.bucket_offsets_per_shard = bucket_offsets_per_shard
        self.bucket_size = bucket_size


def delegating_named_parameters(module: nn.Module, prefix: str = "", recurse: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
    """named_parameters for sharded *wrapper* modules (FP / MC / ITEP): own parameters + each child's own
    ``named_parameters`` (so an inner sharded collection reports its table-keyed names, not its storage layout)."""
    yield from nn.Module.named_parameters(module, prefix, recurse=False)
    if recurse:
        for name, child in module.named_children():
            yield from child.named_parameters((prefix + "." if prefix else "") + name, recurse)


def delegating_state_dict(module: nn.Module, destination: Optional[Dict[str, Any]] = None, prefix: str = "", keep_vars: bool = False) -> Dict[str, Any]:
    """state_dict for sharded *wrapper* modules: own tensors + every child's OWN ``state_dict`` (an inner sharded collection emits
    its table-keyed ShardedTensors instead of its storage layout), so the keys equal the unsharded module's."""
    from collections import OrderedDict

    if destination is None:
        destination = OrderedDict()
    for name, p in module._parameters.items():
        if p is not None:
            destination[prefix + name] = p if keep_vars else p.detach()
    for name, b in module._buffers.items():
        if b is not None and name not in module._non_persistent_buffers_set:
            destination[prefix + name] = b if keep_vars else b.detach()
    for name, child in module.named_children():
        child.state_dict(destination=destination, prefix=prefix + name + ".", keep_vars=keep_vars)
    return destination


def delegating_load_state_dict(module: nn.Module, state_dict: Dict[str, Any], strict: bool = True):
    """Counterpart of :func:`delegating_state_dict`: routes every key prefix to the child's own ``load_state_dict``."""
    from torch.nn.modules.module import _IncompatibleKeys

    missing: List[str] = []
    unexpected: List[str] = []
    own = set()
    with torch.no_grad():
        for name, t in list(module._parameters.items()) + [(n, b) for n, b in module._buffers.items() if n not in module._non_persistent_buffers_set]:
            if t is None:
                continue
            own.add(name)
            if name in state_dict:
                t.copy_(state_dict[name])
            else:
                missing.append(name)
    children = dict(module.named_children())
    for name, child in children.items():
        pre = name + "."
        sub = {k[len(pre):]: v for k, v in state_dict.items() if k.startswith(pre)}
        res = child.load_state_dict(sub, strict=False)
        missing.extend(pre + k for k in res.missing_keys)
        unexpected.extend(pre + k for k in res.unexpected_keys)
    for k in state_dict:
        if k not in own and k.split(".", 1)[0] not in children:
            unexpected.append(k)
    if strict and (missing or unexpected):
        raise RuntimeError(f"Error(s) in loading state_dict for {type(module).__name__}: missing {missing}, unexpected {unexpected}")
    return _IncompatibleKeys(missing, unexpected)


# ---- well-known resources / devices, storage accounting ------------------------------------------------------------------------------------------------
class ParameterStorage(Enum):
    """Physical memories a planner constraint can name."""

    HBM = "hbm"  # GPU-attached
    DDR = "ddr"  # CPU-attached


class StorageUsageType(Enum):
    BASE_QUANT = "BaseQuantEmbeddingSharder"
    BASE = "BaseEmbeddingSharder"
    DEFAULT = "ModuleSharder"


class ComputeDevice(Enum):
    CUDA = "cuda"
    CPU = "cpu"
    MTIA = "mtia"


@unique
class ComputeKernel(Enum):
    DEFAULT = "default"


def compute_storage_usage(tensor: torch.Tensor, compute_device_type: str, compute_kernel: str, storage_usage_type: StorageUsageType) -> Dict[str, int]:
    """{memory: bytes} a parameter needs under a kernel: the tensor's bytes (+ 4 bytes per row of scale / bias for quantized tables); the
    UVM kernels keep the rows in host memory, everything else in the memory of the compute device."""
    size = tensor.element_size() * tensor.nelement()
    if storage_usage_type == StorageUsageType.BASE_QUANT:
        size += tensor.shape[0] * 4
        if compute_kernel in {"quant_uvm", "quant_uvm_caching"}:
            return {ParameterStorage.DDR.value: size}
        where = {"cuda": ParameterStorage.HBM, "cpu": ParameterStorage.DDR, "mtia": ParameterStorage.DDR}
    elif storage_usage_type == StorageUsageType.BASE:
        if compute_kernel in {"fused_uvm", "fused_uvm_caching"}:
            return {ParameterStorage.DDR.value: size}
        where = {"cuda": ParameterStorage.HBM, "cpu": ParameterStorage.DDR, "mtia": ParameterStorage.HBM}
    else:
        where = {"cuda": ParameterStorage.HBM, "cpu": ParameterStorage.DDR, "mtia": ParameterStorage.HBM}
    return {where.get(compute_device_type, ParameterStorage.HBM).value: size}


class DeviceToHostTensorAwaitable(LazyAwaitable[torch.Tensor]):
    """A device tensor on its way to the host: the copy is enqueued at construction (non-blocking), ``wait()`` blocks on the event
    recorded behind it."""

    def __init__(self, tensor_on_device: torch.Tensor) -> None:
        super().__init__()
        self._tensor = tensor_on_device.to("cpu", non_blocking=True)
        self._event = None
        if tensor_on_device.is_cuda:
            self._event = torch.cuda.Event()
            self._event.record()

    def _wait_impl(self) -> torch.Tensor:
        if self._event is not None:
            self._event.synchronize()
        return self._tensor


KT = TypeVar("KT")
VT_co = TypeVar("VT_co")
ParentW = TypeVar("ParentW")


class GetItemLazyAwaitable(LazyAwaitable[W], Generic[W, ParentW, KT]):
    """``parent[key]`` without waiting for the parent yet: waits for it and indexes the result when its own value is asked for."""

    def __init__(self, parent: LazyAwaitable, key: Any) -> None:
        super().__init__()
        self._parent = parent
        self._key = key

    def _wait_impl(self) -> W:
        return LazyAwaitable._wait_async(self._parent)[self._key]


class LazyGetItemMixin(Generic[KT, VT_co]):
    """For LazyAwaitables of mappings: ``awaitable[key]`` returns a ``GetItemLazyAwaitable`` instead of waiting."""

    def __getitem__(self, key: Any) -> "GetItemLazyAwaitable":
        return GetItemLazyAwaitable(self, key)  # type: ignore[arg-type]
