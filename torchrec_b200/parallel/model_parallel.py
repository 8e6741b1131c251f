"""DistributedModelParallel: entry point of model parallelism
(reference torchrec/distributed/model_parallel.py:246-905).

Walks the module tree, replaces every module that has a plan entry by its sharded twin,
materialises remaining meta-device parameters, wraps the dense remainder in DDP (sharded parameters
are ignored by DDP) and exposes the fused optimizers as one ``CombinedOptimizer`` whose state keys are
parameter FQNs — the checkpoint layout of the unsharded model.
"""
from __future__ import annotations

import abc
import copy
from collections import OrderedDict
from typing import Any, Dict, Iterator, List, Mapping, Optional, Set, Tuple, Type, Union, cast

import torch
import torch.distributed as dist
from torch import nn
from torch.nn.modules.module import _IncompatibleKeys
from torch.nn.parallel import DistributedDataParallel

from ..optim.fused import FusedOptimizerModule
from ..optim.keyed import CombinedOptimizer, KeyedOptimizer
from .comm import get_local_size
from .types import EmbeddingModuleShardingPlan, ModuleSharder, ShardedModule, ShardingEnv, ShardingEnv2D, ShardingPlan, DMPCollectionConfig, DMPCollectionContext, ShardingStrategy

_DDP_STATE_DICT_PREFIX = "module."


class DataParallelWrapper(abc.ABC):
    """Wraps the non-sharded part of a DMP model for data parallelism."""

    @abc.abstractmethod
    def wrap(self, dmp: "DistributedModelParallel", env: ShardingEnv, device: torch.device) -> None:
        ...


class DefaultDataParallelWrapper(DataParallelWrapper):
    """DDP over every parameter that was not sharded (reference model_parallel.py:133-213)."""

    def __init__(self, bucket_cap_mb: int = 25, static_graph: bool = True, find_unused_parameters: bool = False,
                 allreduce_comm_precision: Optional[str] = None, params_to_ignore: Optional[List[str]] = None,
                 ddp_kwargs: Optional[Dict[str, Any]] = None) -> None:
        self._bucket_cap_mb = bucket_cap_mb
        self._static_graph = static_graph
        self._find_unused_parameters = find_unused_parameters
        self._allreduce_comm_precision = allreduce_comm_precision
        self._additional_params_to_ignore: Set[str] = set(params_to_ignore or [])
        self._ddp_kwargs: Dict[str, Any] = ddp_kwargs or {}

    def wrap(self, dmp: "DistributedModelParallel", env: ShardingEnv, device: torch.device) -> None:
        pg = env.process_group
        if pg is None:
            raise RuntimeError("Can only init process group with a process group")
        all_names = {key for key, _ in dmp.named_parameters()}
        sharded = set(DistributedModelParallel._sharded_parameter_names(dmp.module))
        params_to_ignore = sharded.union(self._additional_params_to_ignore)
        if params_to_ignore.issuperset(all_names) or not any(p.requires_grad for n, p in dmp.module.named_parameters() if n not in params_to_ignore):
            return  # nothing left for data parallelism
        DistributedDataParallel._set_params_and_buffers_to_ignore_for_model(module=dmp.module, params_and_buffers_to_ignore=params_to_ignore)
        dmp._dmp_wrapped_module = cast(nn.Module, DistributedDataParallel(
            module=dmp._dmp_wrapped_module.to(device),
            device_ids=None if device.type == "cpu" else [device],
            process_group=pg,
            gradient_as_bucket_view=True,
            broadcast_buffers=False,
            static_graph=self._static_graph,
            find_unused_parameters=self._find_unused_parameters,
            bucket_cap_mb=self._bucket_cap_mb,
            **self._ddp_kwargs,
        ))
        if self._allreduce_comm_precision == "fp16":
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks as ddp_default_hooks

            dmp._dmp_wrapped_module.register_comm_hook(None, ddp_default_hooks.fp16_compress_hook)
        elif self._allreduce_comm_precision == "bf16":
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks as ddp_default_hooks

            dmp._dmp_wrapped_module.register_comm_hook(None, ddp_default_hooks.bf16_compress_hook)


def get_unwrapped_module(module: nn.Module) -> nn.Module:
    """Strip DMP / DDP / FSDP wrappers."""
    while isinstance(module, (DistributedModelParallel, DistributedDataParallel)) or type(module).__name__ == "FullyShardedDataParallel":
        if isinstance(module, DistributedModelParallel):
            module = module._dmp_wrapped_module
        else:
            module = module.module
    return module


def get_module(module: nn.Module) -> nn.Module:
    while isinstance(module, DistributedModelParallel):
        module = module._dmp_wrapped_module
    return module


class DistributedModelParallel(nn.Module, FusedOptimizerModule):
    """Entry point to model parallelism.

    Args:
        module: model to wrap (embedding collections may live on the ``meta`` device).
        env: sharding environment (defaults to the world process group).
        device: compute device.
        plan: sharding plan; when ``None`` the planner builds one collectively.
        sharders: module sharders (defaults to ``get_default_sharders()``).
        init_data_parallel / init_parameters / data_parallel_wrapper: as in the reference.

    Example::

        model = DistributedModelParallel(DLRM(ebc, ...), device=torch.device("cuda"))
        dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(model.named_parameters())), lambda p: torch.optim.SGD(p, lr=0.1))
        opt = CombinedOptimizer([model.fused_optimizer, dense_opt])
    """

    def __init__(
        self,
        module: nn.Module,
        env: Optional[ShardingEnv] = None,
        device: Optional[torch.device] = None,
        plan: Optional[ShardingPlan] = None,
        sharders: Optional[List[ModuleSharder[nn.Module]]] = None,
        init_data_parallel: bool = True,
        init_parameters: bool = True,
        data_parallel_wrapper: Optional[DataParallelWrapper] = None,
        model_tracker_config: Optional[Any] = None,
        model_tracker_configs: Optional[Any] = None,
    ) -> None:
        super().__init__()
        torch._C._log_api_usage_once(f"torchrec_b200.parallel.{self.__class__.__name__}")
        self.init_parameters = init_parameters
        self._ddp_wrapped: bool = False
        if env is None:
            pg = dist.GroupMember.WORLD if dist.is_initialized() else None
            env = ShardingEnv.from_process_group(pg) if pg is not None else ShardingEnv.from_local(1, 0)
        self._env: ShardingEnv = env
        self.device: torch.device = torch.device(device) if device is not None else torch.device("cpu")
        if sharders is None:
            from .sharding_plan import get_default_sharders

            sharders = get_default_sharders()
        self._sharder_map: Dict[Type[nn.Module], ModuleSharder[nn.Module]] = {sharder.module_type: sharder for sharder in sharders}
        if data_parallel_wrapper is None:
            data_parallel_wrapper = DefaultDataParallelWrapper()
        self._data_parallel_wrapper: DataParallelWrapper = data_parallel_wrapper
        if plan is None:
            from .planner import EmbeddingShardingPlanner, Topology

            planner = EmbeddingShardingPlanner(topology=Topology(
                local_world_size=get_local_size(self._env.world_size), world_size=self._env.world_size, compute_device=self.device.type))
            pg = self._env.process_group
            plan = planner.collective_plan(module, sharders, pg) if pg is not None else planner.plan(module, sharders)
        self._plan: ShardingPlan = plan
        self._dmp_wrapped_module: nn.Module = self._init_dmp(module)
        self._optim: CombinedOptimizer = self._init_optim(self._dmp_wrapped_module)
        if init_parameters:
            self._init_parameters(self.module)
        if init_data_parallel:
            self.init_data_parallel()
        from ..modules.utils import reset_module_states_post_sharding

        reset_module_states_post_sharding(self._dmp_wrapped_module)  # caches taken on the unsharded model (KTRegroupAsDict, ...) are stale now
        self._model_tracker = None
        if model_tracker_config is not None:
            from .model_tracker import ModelDeltaTracker

            cfg = model_tracker_config
            if isinstance(cfg, dict):
                kw = dict(cfg)
            else:  # DeltaTrackerConfig dataclass
                kw = dict(consumers=getattr(cfg, "consumers", None), delete_on_read=getattr(cfg, "delete_on_read", True), auto_compact=getattr(cfg, "auto_compact", False),
                          mode=getattr(cfg, "tracking_mode", None) or getattr(cfg, "mode"), fqns_to_skip=getattr(cfg, "fqns_to_skip", ()))
            self._model_tracker = ModelDeltaTracker(self._dmp_wrapped_module, **kw)
        # ``ModelTrackerConfigs`` (the reference's newer argument): trackers by kind, kept in ``model_trackers``
        self.model_trackers: Dict[str, Any] = {}
        if model_tracker_configs is not None:
            from .model_tracker import ModelDeltaTracker
            from .model_tracker.trackers.raw_id_tracker import RawIdTracker
            from .model_tracker.types import Trackers

            raw = getattr(model_tracker_configs, "raw_id_tracker_config", None)
            if raw is not None:
                self.model_trackers[Trackers.RAW_ID_TRACKER.name] = RawIdTracker(self._dmp_wrapped_module, delete_on_read=raw.delete_on_read, fqns_to_skip=raw.fqns_to_skip)
            delta = getattr(model_tracker_configs, "delta_tracker_config", None)
            if delta is not None and self._model_tracker is None:
                self._model_tracker = ModelDeltaTracker(self._dmp_wrapped_module, consumers=delta.consumers, delete_on_read=delta.delete_on_read,
                                                        auto_compact=delta.auto_compact, mode=delta.tracking_mode, fqns_to_skip=getattr(delta, "fqns_to_skip", ()))
            if self._model_tracker is not None:
                self.model_trackers[Trackers.DELTA_TRACKER.name] = self._model_tracker

    # ---- public surface -------------------------------------------------------------------------------------
    @property
    def module(self) -> nn.Module:
        """The wrapped model without DDP."""
        return get_unwrapped_module(self)

    @module.setter
    def module(self, value: nn.Module) -> None:
        if isinstance(self.module, DistributedDataParallel):
            raise RuntimeError("module can't be set after calling init_data_parallel(...)")
        self._dmp_wrapped_module = value

    def forward(self, *args, **kwargs) -> Any:
        return self._dmp_wrapped_module(*args, **kwargs)

    def init_data_parallel(self) -> None:
        """Wrap the dense part in DDP (call after meta-device parameters were materialised)."""
        if not self._ddp_wrapped:
            if self._env.process_group is not None and getattr(self._env, "global_world_size", self._env.world_size) > 1:
                self._data_parallel_wrapper.wrap(self, self._env, self.device)
            else:
                self._dmp_wrapped_module = self._dmp_wrapped_module.to(self.device) if self.device.type != "meta" else self._dmp_wrapped_module
            self._ddp_wrapped = True

    def copy(self, device: torch.device) -> "DistributedModelParallel":
        assert isinstance(device, torch.device)
        with torch.no_grad():
            copy_dmp_wrapped_module = copy.deepcopy(self._dmp_wrapped_module).to(device)
        new = copy.copy(self)
        new.device = device
        new._dmp_wrapped_module = copy_dmp_wrapped_module
        return new

    @property
    def plan(self) -> ShardingPlan:
        return self._plan

    @property
    def fused_optimizer(self) -> KeyedOptimizer:
        return self._optim

    # ---- init ------------------------------------------------------------------------------------------------
    def _init_dmp(self, module: nn.Module) -> nn.Module:
        return self._shard_modules_impl(module)

    def _init_optim(self, module: nn.Module) -> CombinedOptimizer:
        return CombinedOptimizer(self._fused_optim_impl(module, []))

    def _fused_optim_impl(self, module: nn.Module, fused_optims: List[Tuple[str, KeyedOptimizer]], path: str = "") -> List[Tuple[str, KeyedOptimizer]]:
        if isinstance(module, FusedOptimizerModule):
            fused_optims.append((path, module.fused_optimizer))
            return fused_optims
        for name, child in module.named_children():
            self._fused_optim_impl(child, fused_optims, path + "." + name if path else name)
        return fused_optims

    def _shard_modules_impl(self, module: nn.Module, path: str = "") -> nn.Module:
        if isinstance(module, ShardedModule):
            return module
        module_sharding_plan = self._plan.get_plan_for_module(path)
        if module_sharding_plan:
            sharder_key = type(module)
            if sharder_key not in self._sharder_map:
                raise RuntimeError(f"no sharder registered for {sharder_key} (plan entry '{path}')")
            module = self._sharder_map[sharder_key].shard(module, module_sharding_plan, self._env, self.device, path)
            torch._C._log_api_usage_once(f"torchrec_b200.parallel.sharded.{type(module).__name__}")
            return module
        for name, child in module.named_children():
            child = self._shard_modules_impl(child, path + "." + name if path else name)
            setattr(module, name, child)
        return module

    def _init_parameters(self, module: nn.Module) -> None:
        @torch.no_grad()
        def init_parameters(m: nn.Module) -> None:
            if isinstance(m, ShardedModule):
                return
            has_meta_param = False
            for name, param in list(m._parameters.items()):
                if isinstance(param, torch.Tensor) and param.device.type == "meta":
                    m._parameters[name] = nn.Parameter(torch.empty_like(param, device=self.device), requires_grad=param.requires_grad)
                    has_meta_param = True
            for name, buffer in list(m._buffers.items()):
                if isinstance(buffer, torch.Tensor) and buffer.device.type == "meta":
                    m._buffers[name] = torch.zeros_like(buffer, device=self.device)
            if has_meta_param and hasattr(m, "reset_parameters"):
                m.reset_parameters()

        def walk(m: nn.Module) -> None:
            if isinstance(m, ShardedModule):
                return
            init_parameters(m)
            for c in m.children():
                walk(c)

        walk(module)

    # ---- state dict / parameters pass-through (keys of the unsharded model) -----------------------------------
    def sparse_grad_parameter_names(self, destination: Optional[List[str]] = None, prefix: str = "") -> List[str]:
        destination = [] if destination is None else destination
        return self._sparse_grad_parameter_names(self.module, destination, prefix)

    def _sparse_grad_parameter_names(self, module: nn.Module, destination: List[str], prefix: str = "") -> List[str]:
        module = get_unwrapped_module(module)
        if isinstance(module, ShardedModule):
            pass
        elif isinstance(module, nn.Embedding):
            if module.sparse:
                destination.append(append_prefix(prefix, "weight"))
        elif isinstance(module, nn.EmbeddingBag):
            if module.sparse:
                destination.append(append_prefix(prefix, "weight"))
        else:
            for name, child in module.named_children():
                self._sparse_grad_parameter_names(child, destination, append_prefix(prefix, name))
        return destination

    def state_dict(self, destination: Optional[Dict[str, Any]] = None, prefix: str = "", keep_vars: bool = False) -> Dict[str, Any]:
        state_dict = get_module(self).state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars)
        torch.nn.modules.utils.consume_prefix_in_state_dict_if_present(state_dict, prefix + _DDP_STATE_DICT_PREFIX)
        return state_dict

    def load_state_dict(self, state_dict: "OrderedDict[str, torch.Tensor]", prefix: str = "", strict: bool = True, assign: bool = False) -> _IncompatibleKeys:
        return self._load_state_dict(self, state_dict, prefix, strict)

    def _load_state_dict(self, module: nn.Module, state_dict: "OrderedDict[str, torch.Tensor]", prefix: str = "", strict: bool = True) -> _IncompatibleKeys:
        missing_keys: List[str] = []
        unexpected_keys: List[str] = []
        module = get_module(module)
        if isinstance(module, DistributedDataParallel):
            module = module.module
        if isinstance(module, ShardedModule):
            sub = OrderedDict((k[len(prefix):], v) for k, v in state_dict.items() if k.startswith(prefix))
            res = module.load_state_dict(sub, strict=strict)
            return _IncompatibleKeys([prefix + k for k in res.missing_keys], [prefix + k for k in res.unexpected_keys])
        # direct params / buffers
        own = OrderedDict()
        for k, v in state_dict.items():
            if k.startswith(prefix) and "." not in k[len(prefix):]:
                own[k[len(prefix):]] = v
        local_names = set(dict(module._parameters).keys()) | {k for k in module._buffers.keys() if k not in module._non_persistent_buffers_set}
        with torch.no_grad():
            for name in local_names:
                t = module._parameters.get(name, None)
                if t is None:
                    t = module._buffers.get(name, None)
                if t is None:
                    continue
                if name in own:
                    t.copy_(own[name])
                else:
                    missing_keys.append(prefix + name)
        for k in own:
            if k not in local_names:
                unexpected_keys.append(prefix + k)
        for name, child in module.named_children():
            res = self._load_state_dict(child, state_dict, prefix + name + ".", strict)
            missing_keys.extend(res.missing_keys)
            unexpected_keys.extend(res.unexpected_keys)
        if prefix == "" and strict:
            known = set()
            for k in state_dict.keys():
                known.add(k)
            if missing_keys:
                raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing_keys}")
        return _IncompatibleKeys(missing_keys, unexpected_keys)

    def _named_parameters(self, module: nn.Module, prefix: str = "", recurse: bool = True, strip_ddp: bool = True) -> Iterator[Tuple[str, torch.nn.Parameter]]:
        if strip_ddp:
            module = get_unwrapped_module(module)
        if isinstance(module, ShardedModule):
            yield from module.named_parameters(prefix, recurse)
        else:
            yield from module.named_parameters(prefix, recurse=False)
            for name, child in module.named_children():
                yield from self._named_parameters(child, append_prefix(prefix, name), recurse, strip_ddp)

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, torch.nn.Parameter]]:
        gen = self._named_parameters(self.module, prefix, recurse)
        memo = set()
        for key, param in gen:
            if param in memo:
                continue
            if remove_duplicate:
                memo.add(param)
            yield key, param

    def bare_named_parameters(self, prefix: str = "", recurse: bool = True) -> Iterator[Tuple[str, torch.nn.Parameter]]:
        gen = self._named_parameters(self.module, prefix, recurse)
        memo = set()
        for key, param in gen:
            if param in memo:
                continue
            memo.add(param)
            yield key, param

    @staticmethod
    def _sharded_parameter_names(module: nn.Module, prefix: str = "") -> Iterator[str]:
        module = get_unwrapped_module(module)
        if isinstance(module, ShardedModule):
            yield from module.sharded_parameter_names(prefix)
        else:
            for name, child in module.named_children():
                yield from DistributedModelParallel._sharded_parameter_names(child, append_prefix(prefix, name))

    def _named_buffers(self, module: nn.Module, prefix: str = "", recurse: bool = True) -> Iterator[Tuple[str, torch.Tensor]]:
        module = get_unwrapped_module(module)
        if isinstance(module, ShardedModule):
            yield from module.named_buffers(prefix, recurse)
        else:
            yield from module.named_buffers(prefix, recurse=False)
            for name, child in module.named_children():
                yield from self._named_buffers(child, append_prefix(prefix, name), recurse)

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, torch.Tensor]]:
        gen = self._named_buffers(self.module, prefix, recurse)
        memo = set()
        for key, param in gen:
            if param in memo:
                continue
            if remove_duplicate:
                memo.add(param)
            yield key, param

    @property
    def fused_optimizer_modules(self) -> List[nn.Module]:
        return [m for m in self.module.modules() if isinstance(m, FusedOptimizerModule)]

    def get_model_tracker(self):
        assert self._model_tracker is not None, "Model tracker is not initialized. Add ModelTrackerConfig at DistributedModelParallel init."
        return self._model_tracker

    def get_delta(self, consumer: Optional[str] = None):
        assert self._model_tracker is not None, "Model tracker is not initialized."
        return self._model_tracker.get_delta(consumer)

    def reshard(self, sharded_module_fqn: str, changed_shard_to_params: Dict[str, Any]) -> None:
        """Move table shards to a new placement at runtime (dynamic re-sharding, reference model_parallel.py:793)."""
        from .sharding.dynamic_sharding import reshard_module

        steps = sharded_module_fqn.split(".")
        parent = self.module
        for s in steps[:-1]:
            parent = getattr(parent, s)
        old = getattr(parent, steps[-1])
        new = reshard_module(old, changed_shard_to_params, self._env, self.device, self._sharder_map)
        setattr(parent, steps[-1], new)
        plan_for = self._plan.plan[sharded_module_fqn]
        for k, v in changed_shard_to_params.items():
            plan_for[k] = v  # type: ignore[index]
        self._optim = self._init_optim(self._dmp_wrapped_module)


def append_prefix(prefix: str, name: str) -> str:
    if prefix != "" and name != "":
        return prefix + "." + name
    return prefix + name


def add_prefix_to_state_dict(state_dict: Dict[str, Any], prefix: str) -> None:
    keys = sorted(state_dict.keys())
    for key in keys:
        state_dict[prefix + key] = state_dict.pop(key)
    if "_metadata" in state_dict:
        metadata = state_dict["_metadata"]
        for key in list(metadata.keys()):
            if len(key) == 0:
                continue
            metadata[prefix + key] = metadata.pop(key)


class DMPCollection(DistributedModelParallel):
    """2D parallelism: model parallel inside *sharding groups* of ``sharding_group_size`` ranks, data parallel across the
    ``world_size // sharding_group_size`` replicas of every shard (reference model_parallel.py:1008-1770).

    The plan is written for ranks ``0..sharding_group_size-1``; every sharding group instantiates it over its own process
    group (the lookup engine only ever sees group-local ranks, so no plan re-mapping is needed). Dense parameters are DDP'd
    over the global group; embedding shards (+ fused optimizer state) are averaged across their replica group by ``sync()``,
    which the training loop calls every N steps.

    ``use_inter_host_allreduce=True`` makes sharding groups contiguous (all-to-all stays inside one NVLink domain, replicas
    talk across hosts) — the natural layout for NVSwitch nodes; the default interleaves them like the reference."""

    def __init__(self, module: nn.Module, device: torch.device, plan: ShardingPlan, world_size: int, sharding_group_size: int,
                 global_pg: dist.ProcessGroup, sharding_strategy: ShardingStrategy = ShardingStrategy.DEFAULT, node_group_size: Optional[int] = None,
                 sharders: Optional[List[ModuleSharder[nn.Module]]] = None, init_data_parallel: bool = True, init_parameters: bool = True,
                 data_parallel_wrapper: Optional[DataParallelWrapper] = None, use_inter_host_allreduce: bool = False,
                 custom_all_reduce: Optional[Any] = None, submodule_configs: Optional[List[DMPCollectionConfig]] = None,
                 rs_awaitable_hook_module: Optional[str] = None) -> None:
        assert world_size % sharding_group_size == 0, "world_size must be a multiple of sharding_group_size"
        self._global_pg_ = global_pg
        self._global_rank = dist.get_rank(global_pg)
        self._custom_all_reduce = custom_all_reduce
        self._world_size_2d = world_size
        self._default_ctx = DMPCollectionContext(module=None, plan=plan, sharding_group_size=sharding_group_size, node_group_size=node_group_size,  # type: ignore[arg-type]
                                                 use_inter_host_allreduce=use_inter_host_allreduce, sharding_strategy=sharding_strategy)
        self._submodule_ctxs = [DMPCollectionContext(module=c.module, plan=c.plan, sharding_group_size=c.sharding_group_size, node_group_size=c.node_group_size,
                                                     use_inter_host_allreduce=c.use_inter_host_allreduce, sharding_strategy=c.sharding_strategy)
                                for c in (submodule_configs or [])]
        self._ctxs: List[DMPCollectionContext] = [self._default_ctx] + self._submodule_ctxs
        for ctx in self._ctxs:
            ctx.device_mesh, ctx.sharding_pg, ctx.replica_pg = self._create_process_groups(self._global_rank, world_size, ctx.sharding_group_size, ctx.use_inter_host_allreduce)
        env = ShardingEnv2D(sharding_pg=self._default_ctx.sharding_pg, global_pg=global_pg, device_mesh=self._default_ctx.device_mesh, node_group_size=node_group_size,
                            use_inter_host_allreduce=use_inter_host_allreduce, replica_pg=self._default_ctx.replica_pg)
        super().__init__(module, env, device, plan, sharders, init_data_parallel, init_parameters, data_parallel_wrapper)
        self._sync_cache: Optional[List[torch.Tensor]] = None
        self._fully_sharded: List[Any] = []
        self._attach_fully_sharded()

    # ---- FULLY_SHARDED strategy (reference batched_embedding_kernel.py:4425-4640, model_parallel.py:1470-1530) --------------
    def _engines(self, ctx: DMPCollectionContext):
        for sharded, _ in ctx.modules_to_sync:
            inner = [m for m in sharded.modules() if hasattr(m, "engine")] or [sharded]
            for m in inner:
                eng = getattr(m, "engine", None)
                if eng is not None:
                    yield eng

    def _attach_fully_sharded(self) -> None:
        """Embedding shards of FULLY_SHARDED contexts are replica-averaged after every forward lookup and held as 1/R slices
        until their backward (``parallel/fully_sharded.py``)."""
        from .fully_sharded import attach

        for ctx in self._ctxs:
            if ctx.sharding_strategy != ShardingStrategy.FULLY_SHARDED or ctx.replica_pg is None or dist.get_world_size(ctx.replica_pg) == 1:
                continue
            for eng in self._engines(ctx):
                for tbe in eng._tbes:
                    fs = attach(tbe, ctx.replica_pg)
                    if fs is not None:
                        self._fully_sharded.append(fs)

    def await_rs_awaitables(self) -> None:
        """Make every FULLY_SHARDED weight buffer whole again (before a checkpoint, an eval pass that bypasses the lookup
        functions, or a plan change). Collective over the replica groups. Reference: model_parallel.py:1472-1503."""
        for fs in self._fully_sharded:
            fs.gather()

    def state_dict(self, *args: Any, **kwargs: Any):  # type: ignore[override]
        self.await_rs_awaitables()
        return super().state_dict(*args, **kwargs)

    # ---- process groups --------------------------------------------------------------------------------------
    @staticmethod
    def _create_process_groups(global_rank: int, world_size: int, local_size: int, use_inter_host_allreduce: bool = False):
        """Returns (mesh as a [replicas, shards] rank matrix, my sharding pg, my replica pg). Collective: every rank
        creates every group in the same order."""
        R = world_size // local_size
        if use_inter_host_allreduce:
            matrix = [[g * local_size + i for i in range(local_size)] for g in range(R)]  # row g = sharding group g (contiguous)
        else:
            matrix = [[g + i * R for i in range(local_size)] for g in range(R)]  # interleaved (reference default)
        sharding_pg = replica_pg = None
        for row in matrix:
            pg = dist.new_group(ranks=row)
            if global_rank in row:
                sharding_pg = pg
        for col in range(local_size):
            ranks = [matrix[g][col] for g in range(R)]
            pg = dist.new_group(ranks=ranks)
            if global_rank in ranks:
                replica_pg = pg
        return matrix, sharding_pg, replica_pg

    # ---- sharding: each context's modules are sharded over that context's group -------------------------------------
    def _ctx_for(self, module: nn.Module) -> DMPCollectionContext:
        for ctx in self._submodule_ctxs:
            if ctx.module is not None and isinstance(module, ctx.module):
                return ctx
        return self._default_ctx

    def _shard_modules_impl(self, module: nn.Module, path: str = "", ctx: Optional[DMPCollectionContext] = None) -> nn.Module:
        if isinstance(module, ShardedModule):
            return module
        if ctx is None or ctx is self._default_ctx:
            ctx = self._ctx_for(module)
        plan = ctx.plan if ctx.plan is not None else self._plan
        module_sharding_plan = plan.get_plan_for_module(path)
        if module_sharding_plan:
            sharder_key = type(module)
            env = ShardingEnv2D(sharding_pg=ctx.sharding_pg, global_pg=self._global_pg_, device_mesh=ctx.device_mesh, node_group_size=ctx.node_group_size,
                                use_inter_host_allreduce=ctx.use_inter_host_allreduce, replica_pg=ctx.replica_pg)
            # sharded modules address ranks inside the sharding group
            env.process_group = ctx.sharding_pg
            env.rank = dist.get_rank(ctx.sharding_pg)
            sharded = self._sharder_map[sharder_key].shard(module, module_sharding_plan, env, self.device, path)
            ctx.modules_to_sync.append((sharded, sharded))
            return sharded
        for name, child in module.named_children():
            setattr(module, name, self._shard_modules_impl(child, path + "." + name if path else name, ctx))
        return module

    # ---- replica synchronisation ----------------------------------------------------------------------------------
    def _sync_tensors(self, include_optimizer_state: bool) -> List[Tuple[torch.Tensor, Any]]:
        out: List[Tuple[torch.Tensor, Any]] = []
        for ctx in self._ctxs:
            for sharded, _ in ctx.modules_to_sync:
                inner = [m for m in sharded.modules() if hasattr(m, "engine")] or [sharded]
                for m in inner:
                    eng = getattr(m, "engine", None)
                    if eng is None:
                        continue
                    for tbe in eng._tbes:
                        if ctx.sharding_strategy != ShardingStrategy.FULLY_SHARDED:  # those are averaged by every forward
                            out.append((tbe.weights.data, ctx.replica_pg))
                        if include_optimizer_state:
                            for st in (tbe.state1, tbe.state2):
                                if st is not None and st.numel():
                                    out.append((st, ctx.replica_pg))
        return out

    @torch.no_grad()
    def sync(self, include_optimizer_state: bool = True) -> None:
        """Average embedding shards (and optimizer state) over the replica groups (reference model_parallel.py:1314-1450)."""
        by_pg: Dict[int, Tuple[Any, List[torch.Tensor]]] = {}
        for t, pg in self._sync_tensors(include_optimizer_state):
            by_pg.setdefault(id(pg), (pg, []))[1].append(t)
        for pg, tensors in by_pg.values():
            if pg is None or dist.get_world_size(pg) == 1:
                continue
            n = dist.get_world_size(pg)
            if self._custom_all_reduce is not None:
                self._custom_all_reduce(tensors)
                continue
            for t in tensors:
                if t.dtype in (torch.float32, torch.float64):
                    dist.all_reduce(t, group=pg)
                    t.div_(n)
                else:  # low-precision tables: reduce in fp32
                    f = t.float()
                    dist.all_reduce(f, group=pg)
                    t.copy_((f / n).to(t.dtype))

    def set_all_reduce_hook(self, reduce_hook: Any) -> None:
        self._custom_all_reduce = reduce_hook

    @property
    def device_mesh(self):
        return self._default_ctx.device_mesh

    @property
    def sharding_pg(self) -> dist.ProcessGroup:
        return self._default_ctx.sharding_pg

    @property
    def replica_pg(self) -> dist.ProcessGroup:
        return self._default_ctx.replica_pg


class HybridEvalDMP(DistributedModelParallel):
    """Eval-only DMP with split placement: embedding shards stay where the plan put them (typically CPU / DDR for tables
    that exceed HBM) while ``.to(device)`` moves only the dense part (reference model_parallel.py:908-1005).
    ``share_embedding_memory(pg)`` de-duplicates CPU-resident table storage across the ranks of one host through POSIX
    shared memory: rank 0 of ``pg`` owns the memory, the others map it."""

    def __init__(self, module: nn.Module, *, init_data_parallel: bool = False, **kwargs: Any) -> None:
        super().__init__(module, init_data_parallel=init_data_parallel, **kwargs)
        self.eval()

    def to(self, *args: Any, **kwargs: Any) -> "HybridEvalDMP":  # type: ignore[override]
        def selective(mod: nn.Module) -> None:
            for key, param in mod._parameters.items():
                if param is not None:
                    mod._parameters[key] = nn.Parameter(param.data.to(*args, **kwargs), requires_grad=param.requires_grad)
            for key, buf in mod._buffers.items():
                if buf is not None:
                    mod._buffers[key] = buf.to(*args, **kwargs)
            for child in mod.children():
                if not isinstance(child, ShardedModule):
                    selective(child)

        selective(self.module)
        return self

    def share_embedding_memory(self, pg: dist.ProcessGroup) -> int:
        """Returns the number of bytes now backed by shared memory."""
        from .collective_utils import create_on_rank_and_share_result

        shared_bytes = 0
        for m in self.module.modules():
            eng = getattr(m, "engine", None)
            if eng is None or not hasattr(eng, "_tbes"):
                continue
            for tbe in eng._tbes:
                w = tbe.weights
                if w.device.type != "cpu" or w.numel() == 0:
                    continue
                src = w.data
                shared = create_on_rank_and_share_result(pg, 0, lambda src=src: src.detach().clone())
                w.data = shared  # parameter views are rebuilt lazily from tbe.weights
                if hasattr(m, "_build_param_views"):
                    m._build_param_views()
                shared_bytes += shared.numel() * shared.element_size()
        return shared_bytes
