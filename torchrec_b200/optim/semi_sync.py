"""Semi-synchronous (DiLoCo-style) optimizer (reference torchrec/optim/semi_sync.py:32): every worker runs a
local optimizer; every ``num_local_steps`` steps the parameter deltas are averaged across workers and applied
by a global (outer) optimizer."""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist

from .keyed import KeyedOptimizer

logger = logging.getLogger(__name__)

SEMI_SYNC_GLOBAL_OPTIM_KEY = "semi_sync_global_optim"
SEMI_SYNC_LOCAL_OPTIM_KEY = "semi_sync_local_optim"
SEMI_SYNC_GLOBAL_STATE_KEY = "semi_sync_global_"
SEMI_SYNC_LOCAL_STATE_KEY = "semi_sync_local_"


class SemisyncOptimizer(KeyedOptimizer):
    def __init__(self, global_params: Optional[Any] = None, optimizer: Optional[KeyedOptimizer] = None, global_optimizer: Optional[KeyedOptimizer] = None,
                 num_local_steps: int = 16, semi_sync_worker_shard_group: Optional[dist.ProcessGroup] = None, offload_global_model: bool = False,
                 non_blocking: bool = False, local_optimizer: Optional[KeyedOptimizer] = None, params: Optional[Any] = None) -> None:
        """``global_params``: the parameters the workers average (``model.parameters()``); ``optimizer``: the local (inner) optimizer
        stepped every step; ``global_optimizer``: the outer optimizer stepped every ``num_local_steps`` steps on the averaged
        pseudo-gradient (global model minus worker model). ``local_optimizer`` / ``params`` are accepted as keyword aliases."""
        local_optimizer = optimizer if optimizer is not None else local_optimizer
        if global_params is None:
            global_params = params
        assert local_optimizer is not None and global_optimizer is not None, "both the local and the global optimizer are required"
        self._global_optimizer = global_optimizer
        self._local_optimizer = local_optimizer
        self._optimizer = local_optimizer
        self._worker_model_params: List[torch.Tensor] = list(global_params) if global_params is not None else [p for g in local_optimizer.param_groups for p in g["params"]]
        self._num_local_steps = num_local_steps
        self._local_step_counter = torch.tensor(0, dtype=torch.int64)
        self._global_step_counter = torch.tensor(0, dtype=torch.int64)
        self._worker_shard_group = semi_sync_worker_shard_group
        self._offload = offload_global_model
        dev = "cpu" if offload_global_model else None
        self._global_params = [p.detach().clone().to(dev) if dev else p.detach().clone() for p in self._worker_model_params]
        merged = dict(local_optimizer.params)
        merged.update(global_optimizer.params)
        super().__init__(merged, dict(local_optimizer.state), list(local_optimizer.param_groups))
        self.defaults["_save_param_groups"] = False

    @property
    def param_groups(self):  # type: ignore[override]
        return list(self._local_optimizer.param_groups)

    @param_groups.setter
    def param_groups(self, v) -> None:
        pass

    def zero_grad(self, set_to_none: bool = False) -> None:
        self._local_optimizer.zero_grad(set_to_none=set_to_none)
        self._global_optimizer.zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure: Any = None) -> None:
        self._local_optimizer.step(closure)
        self._local_step_counter += 1
        if int(self._local_step_counter) % self._num_local_steps == 0:
            self._global_step()

    @torch.no_grad()
    def _global_step(self) -> None:
        W = dist.get_world_size(self._worker_shard_group) if dist.is_initialized() else 1
        gparams = [p for g in self._global_optimizer.param_groups for p in g["params"]]
        for wp, gp, opt_p in zip(self._worker_model_params, self._global_params, gparams if len(gparams) == len(self._worker_model_params) else self._worker_model_params):
            delta = gp.to(wp.device) - wp  # pseudo-gradient: where the global model is minus where the worker went
            if W > 1:
                dist.all_reduce(delta, group=self._worker_shard_group)
                delta /= W
            wp.copy_(gp.to(wp.device))
            opt_p.grad = delta if opt_p is wp else delta.to(opt_p.device)
        self._global_optimizer.step()
        for wp, gp in zip(self._worker_model_params, self._global_params):
            gp.copy_(wp.to(gp.device))
        self._global_step_counter += 1

    def state_dict(self) -> Dict[str, Any]:
        return {"state": {SEMI_SYNC_GLOBAL_OPTIM_KEY: self._global_optimizer.state_dict(), SEMI_SYNC_LOCAL_OPTIM_KEY: self._local_optimizer.state_dict(),
                          "local_step_counter": self._local_step_counter, "global_step_counter": self._global_step_counter}}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        st = state_dict["state"]
        self._global_optimizer.load_state_dict(st[SEMI_SYNC_GLOBAL_OPTIM_KEY])
        self._local_optimizer.load_state_dict(st[SEMI_SYNC_LOCAL_OPTIM_KEY])
        self._local_step_counter = st["local_step_counter"]
        self._global_step_counter = st["global_step_counter"]
