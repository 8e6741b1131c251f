"""Planner / sharding config (reference ``torchrec/distributed/test_utils/sharding_config.py``: ``PlannerConfig`` :66, ``ShardingConfig`` :297)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..planner import EmbeddingShardingPlanner, Topology
from ..planner.types import ParameterConstraints
from ..sharding_plan import get_default_sharders
from ..types import ShardingEnv, ShardingPlan


def _detect_hbm_cap(compute_device: str) -> Optional[int]:
    if compute_device == "cuda" and torch.cuda.is_available():
        return int(torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory)
    return None


@dataclass
class PlannerConfig:
    planner_type: str = "embedding"
    world_size: int = 1
    compute_device: str = "cuda"
    local_world_size: Optional[int] = None
    hbm_cap: Optional[int] = None
    batch_size: int = 512
    storage_reservation_percentage: float = 0.15
    constraints: Dict[str, Dict[str, Any]] = field(default_factory=dict)  # table -> ParameterConstraints kwargs

    def generate_topology(self, device_type: Optional[str] = None) -> Topology:
        dev = device_type or self.compute_device
        kw: Dict[str, Any] = dict(world_size=self.world_size, compute_device=dev)
        if self.local_world_size:
            kw["local_world_size"] = self.local_world_size
        cap = self.hbm_cap or _detect_hbm_cap(dev)
        if cap:
            kw["hbm_cap"] = cap
        return Topology(**kw)

    def generate_planner(self, tables: Optional[List[Any]] = None) -> EmbeddingShardingPlanner:
        from ..planner.storage_reservations import HeuristicalStorageReservation

        cons = {name: ParameterConstraints(**kw) for name, kw in self.constraints.items()} or None
        return EmbeddingShardingPlanner(topology=self.generate_topology(), batch_size=self.batch_size, constraints=cons,
                                        storage_reservation=HeuristicalStorageReservation(percentage=self.storage_reservation_percentage))


def _get_sharders_with_fused_params(fused_params: Optional[Dict[str, Any]]) -> List[Any]:
    if not fused_params:
        return get_default_sharders()
    out = []
    for s in get_default_sharders():
        try:
            out.append(type(s)(fused_params=dict(fused_params)))
        except TypeError:
            out.append(s)
    return out


@dataclass
class ShardingConfig:
    """plan + sharders + env -> ``DistributedModelParallel`` and its optimizer."""

    planner: PlannerConfig = field(default_factory=PlannerConfig)
    fused_params: Dict[str, Any] = field(default_factory=dict)
    dense_optimizer: str = "SGD"
    dense_lr: float = 0.1
    sparse_optimizer: str = "EXACT_ROWWISE_ADAGRAD"
    sparse_lr: float = 0.1

    def generate_sharded_model_and_optimizer(self, model: nn.Module, pg: Optional[dist.ProcessGroup], device: torch.device, plan: Optional[ShardingPlan] = None) -> Tuple[nn.Module, torch.optim.Optimizer]:
        from ...optim.keyed import CombinedOptimizer, KeyedOptimizerWrapper
        from ..model_parallel import DistributedModelParallel

        fp = {"optimizer": self.sparse_optimizer, "learning_rate": self.sparse_lr, **self.fused_params}
        sharders = _get_sharders_with_fused_params(fp)
        env = ShardingEnv.from_process_group(pg) if pg is not None else ShardingEnv.from_local(1, 0)
        if plan is None:
            planner = self.planner.generate_planner()
            plan = planner.collective_plan(model, sharders, pg) if pg is not None else planner.plan(model, sharders)
        dmp = DistributedModelParallel(module=model, env=env, device=device, plan=plan, sharders=sharders)
        dense_params = {k: v for k, v in dmp.named_parameters() if v.requires_grad and not hasattr(v, "_in_backward_optimizers")}
        opt_cls = getattr(torch.optim, self.dense_optimizer)
        dense_opt = KeyedOptimizerWrapper(dense_params, lambda params: opt_cls(params, lr=self.dense_lr)) if dense_params else None
        optimizer = CombinedOptimizer([dmp.fused_optimizer] + ([dense_opt] if dense_opt is not None else []))
        return dmp, optimizer
