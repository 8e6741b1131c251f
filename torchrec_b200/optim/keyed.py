"""Keyed optimizers: optimizer state addressed by parameter FQN so checkpoints survive re-sharding
(reference torchrec/optim/keyed.py:34-514)."""
from __future__ import annotations

import json
from copy import deepcopy
from typing import Any, Callable, Collection, Dict, List, Mapping, Optional, OrderedDict, Set, Tuple, Union

import torch
from torch import optim
from torch.distributed._shard.sharded_tensor import ShardedTensor

OptimizerFactory = Callable[[List[Union[torch.Tensor, ShardedTensor]]], optim.Optimizer]


class KeyedOptimizer(optim.Optimizer):
    """Optimizer whose ``state_dict()["state"]`` is keyed by parameter name.

    ``params``: name -> tensor, ``state``: tensor -> state dict, ``param_groups``: usual list.
    ``init_state`` runs one zero-gradient step to materialise lazily created state."""

    def __init__(self, params: Mapping[str, Union[torch.Tensor, ShardedTensor]], state: Mapping[Any, Any], param_groups: Collection[Mapping[str, Any]]) -> None:
        torch._C._log_api_usage_once(f"torchrec_b200.optim.{self.__class__.__name__}")
        self._optimizer_step_pre_hooks: Dict[int, Callable] = {}
        self._optimizer_step_post_hooks: Dict[int, Callable] = {}
        self.state: Mapping[Any, Any] = state
        self.param_groups: Collection[Mapping[str, Any]] = param_groups
        self.params = params
        self.defaults: Dict[str, Any] = {"_save_param_groups": False}
        params_set = set(params.values())
        non_param_state_keys = [p for p in self.state.keys() if p not in params_set]
        if len(non_param_state_keys) > 0:
            raise ValueError("All state keys must be params. The following keys are not: {}.".format(", ".join(str(k) for k in non_param_state_keys)))

    @staticmethod
    def _extract_state_dict_content(input_dict: Dict[str, Any]) -> Dict[str, Any]:
        result: Dict[str, Any] = {}
        for k, v in input_dict.items():
            if isinstance(v, dict):
                result[k] = KeyedOptimizer._extract_state_dict_content(v)
            elif hasattr(v, "state_dict") and callable(v.state_dict):
                result[k] = v.state_dict()
            else:
                result[k] = v
        return result

    @staticmethod
    def _update_param_state_dict_object(current_param_state_dict: Dict[str, Any], param_state_dict_to_load: Dict[str, Any], parent_keys: List[Union[str, int, float, bool, None]]) -> None:
        for k, v in current_param_state_dict.items():
            new_v = param_state_dict_to_load[k]
            parent_keys.append(k)
            if isinstance(v, dict):
                KeyedOptimizer._update_param_state_dict_object(v, new_v, parent_keys)
            elif hasattr(v, "load_state_dict") and callable(v.load_state_dict):
                v.load_state_dict(new_v)
            elif isinstance(v, ShardedTensor):
                assert isinstance(new_v, ShardedTensor)
                num_shards, num_new = len(v.local_shards()), len(new_v.local_shards())
                if num_shards != num_new:
                    raise ValueError(f"Different number of shards {num_shards} vs {num_new} for the path of {json.dumps(parent_keys)}")
                for shard, new_shard in zip(v.local_shards(), new_v.local_shards()):
                    shard.tensor.detach().copy_(new_shard.tensor)
            elif isinstance(v, torch.Tensor):
                v.detach().copy_(new_v)
            else:
                current_param_state_dict[k] = deepcopy(new_v)
            parent_keys.pop()

    def state_dict(self) -> Dict[str, Any]:
        param_groups = self.param_groups
        params = self.params
        param_to_key = {param: key for key, param in params.items()}
        ret_state = {param_to_key[param]: self._extract_state_dict_content(param_state) for param, param_state in self.state.items()}
        ret_groups = []
        for group in param_groups:
            param_keys = [param_to_key[param] for param in group["params"]]
            ret_group = {"params": sorted(param_keys)}
            for k, v in group.items():
                if k != "params":
                    ret_group[k] = deepcopy(v)
            ret_groups.append(ret_group)
        ret: Dict[str, object] = {"state": ret_state}
        if self.defaults["_save_param_groups"]:
            ret["param_groups"] = ret_groups
        return ret

    def post_load_state_dict(self) -> None:
        pass

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        new_state = state_dict["state"]
        state = self.state
        params = self.params
        if len(state) != len(new_state):
            raise ValueError(f"Different parameter count: {len(state)} vs {len(new_state)}")
        for param_key, param in params.items():
            if param not in state:
                continue
            if param_key not in new_state:
                raise ValueError(f"Parameter {param_key} not found")
            if len(state[param]) != len(new_state[param_key]):
                raise ValueError(f"Different state size: {len(state[param])} vs {len(new_state[param_key])}")
            KeyedOptimizer._update_param_state_dict_object(state[param], new_state[param_key], [param_key])
        if self.defaults["_save_param_groups"]:
            new_param_groups = state_dict["param_groups"]
            param_groups = self.param_groups
            if len(param_groups) != len(new_param_groups):
                raise ValueError(f"Different param_groups count: {len(param_groups)} vs {len(new_param_groups)}")
            param_to_key = {param: key for key, param in params.items()}
            group_map = {}
            for group in param_groups:
                group_map["/".join(sorted(param_to_key[param] for param in group["params"]))] = group
            new_group_map = {"/".join(sorted(g["params"])): g for g in new_param_groups}
            for group_key, group in group_map.items():
                if group_key not in new_group_map:
                    raise ValueError(f"Group {group_key} not found")
                new_group = new_group_map[group_key]
                if len(group) != len(new_group):
                    raise ValueError(f"Different param_group size: {len(group)} vs {len(new_group)}")
                for k in group:
                    if k not in new_group:
                        raise ValueError(f"Group key {k} not found for group {group_key}")
                    if k != "params":
                        group[k] = deepcopy(new_group[k])
        self.post_load_state_dict()

    def add_param_group(self, param_group: Any) -> None:
        raise NotImplementedError()

    def init_state(self, sparse_grad_parameter_names: Optional[Set[str]] = None) -> None:
        """Run a step with zero gradients so that lazily-initialised optimizer state exists."""
        for key, param in self.params.items():
            if param.requires_grad:
                t = torch.zeros_like(param)
                if sparse_grad_parameter_names is not None and key in sparse_grad_parameter_names:
                    t = t.to_sparse()
                param.grad = torch.autograd.Variable(t)
        self.step(closure=None)

    def save_param_groups(self, save: bool) -> None:
        self.defaults["_save_param_groups"] = save

    def __getstate__(self) -> Dict[str, object]:
        return self.__dict__


class CombinedOptimizer(KeyedOptimizer):
    """Several KeyedOptimizers behind one interface; keys are prefixed with the optimizer's name."""

    def __init__(self, optims: List[Union[KeyedOptimizer, Tuple[str, KeyedOptimizer]]]) -> None:
        self.defaults: Dict[str, Any] = {}
        self._optims: List[Tuple[str, KeyedOptimizer]] = []
        for key_value in optims:
            if isinstance(key_value, KeyedOptimizer):
                key_value = ("", key_value)
            self._optims.append(key_value)
        all_keys: Set[str] = set()
        self.defaults["_save_param_groups"] = False if len(self._optims) == 0 else self._optims[0][1].defaults["_save_param_groups"]
        for opt_key, opt in self._optims:
            assert self.defaults["_save_param_groups"] == opt.defaults["_save_param_groups"]
            for param_key in opt.params.keys():
                new_param = CombinedOptimizer.prepend_opt_key(param_key, opt_key)
                if new_param in all_keys:
                    raise ValueError(f"Duplicate param key {new_param}")
                all_keys.add(new_param)
        self._optimizer_step_pre_hooks: Dict[int, Callable] = {}
        self._optimizer_step_post_hooks: Dict[int, Callable] = {}
        self._patch_step_function()

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}: {[opt for _, opt in self._optims]}"

    def zero_grad(self, set_to_none: bool = True) -> None:
        for _, opt in self._optims:
            opt.zero_grad(set_to_none=set_to_none)

    def step(self, closure: Any = None) -> None:
        for _, opt in self._optims:
            opt.step(closure=closure)

    @property
    def optimizers(self) -> List[Tuple[str, KeyedOptimizer]]:
        return self._optims

    @staticmethod
    def prepend_opt_key(name: str, opt_key: str) -> str:
        if not name:
            return opt_key
        return opt_key + ("." if opt_key else "") + name

    @property
    def param_groups(self) -> Collection[Mapping[str, Any]]:
        return [pg for _, opt in self._optims for pg in opt.param_groups]

    @property
    def params(self) -> Mapping[str, Union[torch.Tensor, ShardedTensor]]:
        ret = {}
        for opt_key, opt in self._optims:
            for param_key, param in opt.params.items():
                ret[CombinedOptimizer.prepend_opt_key(param_key, opt_key)] = param
        return ret

    @property
    def state(self) -> Mapping[torch.Tensor, Any]:
        ret = {}
        for _, opt in self._optims:
            for param, state in opt.state.items():
                ret[param] = state
        return ret

    def post_load_state_dict(self) -> None:
        for _, opt in self._optims:
            opt.post_load_state_dict()

    def save_param_groups(self, save: bool) -> None:
        self.defaults["_save_param_groups"] = save
        for _, opt in self._optims:
            opt.save_param_groups(save)

    def set_optimizer_step(self, step: int) -> None:
        for _, opt in self._optims:
            if hasattr(opt, "set_optimizer_step"):
                opt.set_optimizer_step(step)


class KeyedOptimizerWrapper(KeyedOptimizer):
    """Wrap a torch optimizer factory into a KeyedOptimizer."""

    def __init__(self, params: Mapping[str, Union[torch.Tensor, ShardedTensor]], optim_factory: OptimizerFactory) -> None:
        self._optimizer: optim.Optimizer = optim_factory(list(params.values()))
        super().__init__(params, self._optimizer.state, self._optimizer.param_groups)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._optimizer.zero_grad(set_to_none=set_to_none)

    def step(self, closure: Any = None) -> None:
        self._optimizer.step(closure=closure)


class OptimizerWrapper(KeyedOptimizer):
    """Base for optimizers that decorate another KeyedOptimizer (clipping, warmup...)."""

    def __init__(self, optimizer: KeyedOptimizer) -> None:
        self._optimizer = optimizer
        self.params: Mapping[str, Union[torch.Tensor, ShardedTensor]] = optimizer.params
        self.state: Mapping[Any, Any] = optimizer.state
        self.param_groups: Collection[Mapping[str, Any]] = optimizer.param_groups
        self.defaults: Dict[str, Any] = {"_save_param_groups": False}
        self._optimizer_step_pre_hooks: Dict[int, Callable] = {}
        self._optimizer_step_post_hooks: Dict[int, Callable] = {}

    def __repr__(self) -> str:
        return self._optimizer.__repr__()

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._optimizer.zero_grad(set_to_none=set_to_none)

    def step(self, closure: Any = None) -> None:
        self._optimizer.step(closure=closure)

    def add_param_group(self, param_group: Any) -> None:
        raise NotImplementedError()

    def state_dict(self) -> Dict[str, Any]:
        return self._optimizer.state_dict()

    def post_load_state_dict(self) -> None:
        self._optimizer.post_load_state_dict()

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        self._optimizer.load_state_dict(state_dict)
        self.state = self._optimizer.state
        self.param_groups = self._optimizer.param_groups
        self.post_load_state_dict()

    def save_param_groups(self, save: bool) -> None:
        self._optimizer.save_param_groups(save)
