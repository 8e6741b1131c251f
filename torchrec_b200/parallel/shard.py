"""Composable sharding API without the DMP wrapper (reference torchrec/distributed/shard.py:55-320)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Type, Union

import torch
import torch.distributed as dist
from torch import nn

from .comm import get_local_size
from .sharding_plan import get_default_sharders, get_module_to_default_sharders
from .types import EmbeddingModuleShardingPlan, ModuleSharder, ShardedModule, ShardingEnv, ShardingPlan, ShardingPlanner


def _join_module_path(path: str, name: str) -> str:
    return (path + "." + name) if path else name


def shard(module: nn.Module, plan: Union[EmbeddingModuleShardingPlan, Callable], env: Optional[Union[ShardingEnv, dist.ProcessGroup]] = None,
          device: Optional[torch.device] = None, sharder: Optional[ModuleSharder[nn.Module]] = None) -> nn.Module:
    """Shard ONE module (e.g. an EmbeddingBagCollection) according to a module sharding plan."""
    if sharder is None:
        sharder = get_module_to_default_sharders().get(type(module), None)
    assert sharder is not None, f"Could not find a valid sharder type for {type(module)}"
    if env is None:
        pg = dist.GroupMember.WORLD if dist.is_initialized() else None
        env = ShardingEnv.from_process_group(pg) if pg is not None else ShardingEnv.from_local(1, 0)
    elif isinstance(env, dist.ProcessGroup):
        env = ShardingEnv.from_process_group(env)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    if callable(plan) and not isinstance(plan, dict):
        from .sharding_plan import apply_to_all, construct_module_sharding_plan

        plan = construct_module_sharding_plan(module, apply_to_all(module, plan, sharder), sharder, world_size=env.world_size,
                                              local_size=get_local_size(env.world_size), device_type=device.type)
    return sharder.shard(module, plan, env, device)


def shard_modules(module: nn.Module, env: Optional[ShardingEnv] = None, device: Optional[torch.device] = None, plan: Optional[ShardingPlan] = None,
                  sharders: Optional[List[ModuleSharder[nn.Module]]] = None, init_params: bool = False, planner: Optional[ShardingPlanner] = None) -> nn.Module:
    """Replace every sub-module that has a plan entry by its sharded twin, in place, and return ``module``."""
    return _shard_modules(module, env, device, plan, sharders, init_params, planner)


def _shard_modules(module: nn.Module, env: Optional[ShardingEnv] = None, device: Optional[torch.device] = None, plan: Optional[ShardingPlan] = None,
                   sharders: Optional[List[ModuleSharder[nn.Module]]] = None, init_params: Optional[bool] = False,
                   planner: Optional[ShardingPlanner] = None) -> nn.Module:
    if sharders is None:
        sharders = get_default_sharders()
    if env is None:
        pg = dist.GroupMember.WORLD if dist.is_initialized() else None
        env = ShardingEnv.from_process_group(pg) if pg is not None else ShardingEnv.from_local(1, 0)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    sharder_map: Dict[Type[nn.Module], ModuleSharder[nn.Module]] = {s.module_type: s for s in sharders}
    if plan is None:
        if planner is None:
            from .planner import EmbeddingShardingPlanner, Topology

            planner = EmbeddingShardingPlanner(topology=Topology(local_world_size=get_local_size(env.world_size), world_size=env.world_size,
                                                                  compute_device=device.type if device.type != "meta" else "cuda"))
        pg = env.process_group
        plan = planner.collective_plan(module, sharders, pg) if pg is not None else planner.plan(module, sharders)

    if type(module) in sharder_map and plan.get_plan_for_module("") is not None:
        return sharder_map[type(module)].shard(module, plan.get_plan_for_module(""), env, device, "")

    def replace(m: nn.Module, path: str) -> None:
        for name, child in list(m.named_children()):
            child_path = _join_module_path(path, name)
            if isinstance(child, ShardedModule):
                continue
            mplan = plan.get_plan_for_module(child_path)
            if mplan is not None and type(child) in sharder_map:
                setattr(m, name, sharder_map[type(child)].shard(child, mplan, env, device, child_path))
            else:
                replace(child, child_path)

    replace(module, "")
    if init_params and device.type != "meta":
        for m in module.modules():
            if isinstance(m, ShardedModule):
                continue
            for n, p in list(m._parameters.items()):
                if p is not None and p.device.type == "meta":
                    m._parameters[n] = nn.Parameter(torch.empty_like(p, device=device), requires_grad=p.requires_grad)
                    if hasattr(m, "reset_parameters"):
                        m.reset_parameters()
    return module
