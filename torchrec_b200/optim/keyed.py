"""Optimizers whose state is addressed by parameter *name*.

A torch optimizer keys its state by parameter object and its checkpoint by position in ``param_groups``; neither survives
re-sharding an embedding table over a different number of ranks. Here every optimizer carries a ``name -> parameter`` map and
its ``state_dict()["state"]`` is ``{name: {...}}`` - sharded state (``ShardedTensor`` / ``DTensor`` over local shards) is
exported as is and loaded shard by shard, in place, so that fused kernels keep pointing at the same storage.

Building blocks
  ``export_state`` / ``import_state``   walk a nested state tree; one handler per leaf kind (sharded tensor, DTensor, tensor,
                                         stateful object, plain value)
  ``KeyedOptimizer``                     the name-keyed view over (params, state, param_groups)
  ``CombinedOptimizer``                  several keyed optimizers behind one interface, names prefixed per optimizer
  ``KeyedOptimizerWrapper``              torch optimizer factory -> keyed optimizer
  ``OptimizerWrapper``                   base of decorators (clipping, warm-up) around a keyed optimizer

Capability parity: reference ``torchrec/optim/keyed.py`` (``KeyedOptimizer`` ``:34``, DTensor branch ``:139-165``,
``CombinedOptimizer`` ``:317``, ``KeyedOptimizerWrapper`` ``:428``, ``OptimizerWrapper`` ``:455``).
"""
from __future__ import annotations

import copy
from typing import Any, Callable, Collection, Dict, Iterable, List, Mapping, Optional, Sequence, Set, Tuple, Union

import torch
from torch import optim
from torch.distributed._shard.sharded_tensor import ShardedTensor

try:  # DTensor layout of sharded state (env.output_dtensor)
    from torch.distributed.tensor import DTensor
except Exception:  # pragma: no cover - very old torch
    DTensor = None  # type: ignore[assignment,misc]

ParamLike = Union[torch.Tensor, ShardedTensor]
OptimizerFactory = Callable[[List[ParamLike]], optim.Optimizer]


class StateMismatch(ValueError):
    """The state being loaded does not have the shape of the optimizer's current state."""


def _where(path: Sequence[Any]) -> str:
    return "/".join(str(p) for p in path) or "<root>"


def _is_stateful(x: Any) -> bool:
    return callable(getattr(x, "state_dict", None)) and callable(getattr(x, "load_state_dict", None))


def export_state(node: Any) -> Any:
    """Checkpoint view of a state tree: dicts are rebuilt, stateful leaves are replaced by their own ``state_dict()``, tensors
    (plain, sharded, DTensor) are referenced, not copied."""
    if isinstance(node, dict):
        return {k: export_state(v) for k, v in node.items()}
    if not isinstance(node, torch.Tensor) and _is_stateful(node):
        return node.state_dict()
    return node


def _local_pieces(t: Any) -> Optional[List[torch.Tensor]]:
    """The rank-local tensors behind a sharded container, or None for anything else."""
    if isinstance(t, ShardedTensor):
        return [s.tensor for s in t.local_shards()]
    if DTensor is not None and isinstance(t, DTensor):
        local = t.to_local()
        shards = getattr(local, "local_shards", None)  # LocalShardsWrapper: several shards of one table on this rank
        return list(shards()) if callable(shards) else [local]
    return None


def import_state(current: Dict[Any, Any], incoming: Mapping[Any, Any], path: List[Any]) -> None:
    """Load ``incoming`` into ``current`` IN PLACE (same keys required): tensors are overwritten shard by shard, stateful objects
    get ``load_state_dict``, plain values are replaced."""
    for key, old in current.items():
        if key not in incoming:
            raise StateMismatch(f"{_where(path + [key])}: missing from the state being loaded")
        new = incoming[key]
        here = path + [key]
        if isinstance(old, dict):
            import_state(old, new, here)
            continue
        pieces = _local_pieces(old)
        if pieces is not None:
            new_pieces = _local_pieces(new)
            if new_pieces is None:
                raise StateMismatch(f"{_where(here)}: expected {type(old).__name__}, got {type(new).__name__}")
            if len(pieces) != len(new_pieces):
                raise StateMismatch(f"{_where(here)}: {len(pieces)} local shard(s) here, {len(new_pieces)} in the state being loaded")
            for dst, src in zip(pieces, new_pieces):
                dst.detach().copy_(src)
        elif isinstance(old, torch.Tensor):
            old.detach().copy_(new)
        elif _is_stateful(old):
            old.load_state_dict(new)
        else:
            current[key] = copy.deepcopy(new)


class KeyedOptimizer(optim.Optimizer):
    """Name-keyed view over an optimizer's (params, state, param_groups).

    ``params``        name -> parameter (plain tensor, ShardedTensor, DTensor)
    ``state``         parameter -> per-parameter state tree (may be shared with a wrapped torch optimizer)
    ``param_groups``  the usual list; saved only after ``save_param_groups(True)``
    """

    def __init__(self, params: Mapping[str, ParamLike], state: Mapping[Any, Any], param_groups: Collection[Mapping[str, Any]]) -> None:
        torch._C._log_api_usage_once(f"torchrec_b200.optim.{type(self).__name__}")
        # torch.optim.Optimizer.__init__ is bypassed on purpose (it would re-group the parameters): provide what step hooks need
        self._optimizer_step_pre_hooks: Dict[int, Callable] = {}
        self._optimizer_step_post_hooks: Dict[int, Callable] = {}
        self.params = params
        self.state: Mapping[Any, Any] = state
        self.param_groups: Collection[Mapping[str, Any]] = param_groups
        self.defaults: Dict[str, Any] = {"_save_param_groups": False}
        owned = set(params.values())
        strays = [k for k in state.keys() if k not in owned]
        if strays:
            raise ValueError("All state keys must be params. The following keys are not: " + ", ".join(str(k) for k in strays) + ".")

    # ---- naming ----------------------------------------------------------------------------------------------
    def _names(self) -> Dict[Any, str]:
        """parameter -> name (rebuilt on demand: decorators may register extra pseudo parameters later)."""
        return {p: n for n, p in self.params.items()}

    @staticmethod
    def _group_id(names: Iterable[str]) -> str:
        return "/".join(sorted(names))

    # ---- checkpoint ------------------------------------------------------------------------------------------
    def state_dict(self) -> Dict[str, Any]:
        names = self._names()
        out: Dict[str, Any] = {"state": {names[p]: export_state(st) for p, st in self.state.items()}}
        if self.defaults["_save_param_groups"]:
            groups = []
            for g in self.param_groups:
                entry = {k: copy.deepcopy(v) for k, v in g.items() if k != "params"}
                entry["params"] = sorted(names[p] for p in g["params"])
                groups.append(entry)
            out["param_groups"] = groups
        return out

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        incoming = state_dict["state"]
        if len(incoming) != len(self.state):
            raise StateMismatch(f"Different parameter count: {len(self.state)} vs {len(incoming)}")
        for name, p in self.params.items():
            mine = self.state.get(p) if hasattr(self.state, "get") else (self.state[p] if p in self.state else None)
            if mine is None:
                continue  # stateless parameter (e.g. never stepped)
            if name not in incoming:
                raise StateMismatch(f"Parameter {name} not found")
            if len(mine) != len(incoming[name]):
                raise StateMismatch(f"Different state size: {len(mine)} vs {len(incoming[name])}")
            import_state(mine, incoming[name], [name])
        if self.defaults["_save_param_groups"]:
            self._load_param_groups(state_dict["param_groups"])
        self.post_load_state_dict()

    def _load_param_groups(self, incoming: Sequence[Mapping[str, Any]]) -> None:
        if len(incoming) != len(self.param_groups):
            raise StateMismatch(f"Different param_groups count: {len(self.param_groups)} vs {len(incoming)}")
        names = self._names()
        by_id = {self._group_id(g["params"]): g for g in incoming}
        for g in self.param_groups:
            gid = self._group_id(names[p] for p in g["params"])
            src = by_id.get(gid)
            if src is None:
                raise StateMismatch(f"Group {gid} not found")
            if len(src) != len(g):
                raise StateMismatch(f"Different param_group size: {len(g)} vs {len(src)}")
            for k in g:
                if k not in src:
                    raise StateMismatch(f"Group key {k} not found for group {gid}")
                if k != "params":
                    g[k] = copy.deepcopy(src[k])  # type: ignore[index]

    def post_load_state_dict(self) -> None:
        """Hook for decorators that derive values (learning rate, counters) from loaded state."""

    def save_param_groups(self, save: bool) -> None:
        self.defaults["_save_param_groups"] = save

    # ---- misc --------------------------------------------------------------------------------------------------
    def add_param_group(self, param_group: Any) -> None:
        raise NotImplementedError("keyed optimizers are built over a fixed parameter map")

    def init_state(self, sparse_grad_parameter_names: Optional[Set[str]] = None) -> None:
        """Materialise lazily created optimizer state by stepping once with all-zero gradients (so that a checkpoint can be
        loaded before the first real step)."""
        sparse = sparse_grad_parameter_names or set()
        for name, p in self.params.items():
            if not p.requires_grad:
                continue
            zero = torch.zeros_like(p)
            p.grad = zero.to_sparse() if name in sparse else zero
        self.step(closure=None)

    def __getstate__(self) -> Dict[str, object]:
        return self.__dict__


class CombinedOptimizer(KeyedOptimizer):
    """Several keyed optimizers stepped, zeroed, saved and loaded as one. ``optims`` entries are optimizers or ``(prefix,
    optimizer)`` pairs; parameter names are exposed as ``prefix.name`` and must be unique across members."""

    def __init__(self, optims: List[Union[KeyedOptimizer, Tuple[str, KeyedOptimizer]]]) -> None:
        self._optims: List[Tuple[str, KeyedOptimizer]] = [o if isinstance(o, tuple) else ("", o) for o in optims]
        save = self._optims[0][1].defaults["_save_param_groups"] if self._optims else False
        if any(o.defaults["_save_param_groups"] != save for _, o in self._optims):
            raise AssertionError("all members of a CombinedOptimizer must agree on save_param_groups")
        self.defaults: Dict[str, Any] = {"_save_param_groups": save}
        seen: Set[str] = set()
        for prefix, o in self._optims:
            for name in o.params.keys():
                full = self.prepend_opt_key(name, prefix)
                if full in seen:
                    raise ValueError(f"Duplicate param key {full}")
                seen.add(full)
        self._optimizer_step_pre_hooks: Dict[int, Callable] = {}
        self._optimizer_step_post_hooks: Dict[int, Callable] = {}
        self._patch_step_function()

    def __repr__(self) -> str:
        return f"{type(self).__name__}: {[o for _, o in self._optims]}"

    @staticmethod
    def prepend_opt_key(name: str, opt_key: str) -> str:
        if not name:
            return opt_key
        return f"{opt_key}.{name}" if opt_key else name

    @property
    def optimizers(self) -> List[Tuple[str, KeyedOptimizer]]:
        return self._optims

    # the three views are unions over the members, recomputed on access (members may grow, e.g. warm-up pseudo parameters)
    @property
    def params(self) -> Mapping[str, ParamLike]:  # type: ignore[override]
        return {self.prepend_opt_key(n, prefix): p for prefix, o in self._optims for n, p in o.params.items()}

    @property
    def state(self) -> Mapping[Any, Any]:  # type: ignore[override]
        merged: Dict[Any, Any] = {}
        for _, o in self._optims:
            merged.update(o.state)
        return merged

    @property
    def param_groups(self) -> Collection[Mapping[str, Any]]:  # type: ignore[override]
        return [g for _, o in self._optims for g in o.param_groups]

    def zero_grad(self, set_to_none: bool = True) -> None:
        for _, o in self._optims:
            o.zero_grad(set_to_none=set_to_none)

    def step(self, closure: Any = None) -> None:
        for _, o in self._optims:
            o.step(closure=closure)

    def post_load_state_dict(self) -> None:
        for _, o in self._optims:
            o.post_load_state_dict()

    def save_param_groups(self, save: bool) -> None:
        self.defaults["_save_param_groups"] = save
        for _, o in self._optims:
            o.save_param_groups(save)

    def set_optimizer_step(self, step: int) -> None:
        for _, o in self._optims:
            if hasattr(o, "set_optimizer_step"):
                o.set_optimizer_step(step)


class KeyedOptimizerWrapper(KeyedOptimizer):
    """``optim_factory(list of parameters) -> torch optimizer`` turned into a keyed optimizer (state and groups are shared with it)."""

    def __init__(self, params: Mapping[str, ParamLike], optim_factory: OptimizerFactory, pass_params_dict: bool = False) -> None:
        # ``pass_params_dict``: the factory wants the name -> parameter mapping (e.g. to build per-name groups) instead of the list
        self._optimizer: optim.Optimizer = optim_factory(params if pass_params_dict else list(params.values()))
        super().__init__(params, self._optimizer.state, self._optimizer.param_groups)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._optimizer.zero_grad(set_to_none=set_to_none)

    def step(self, closure: Any = None) -> None:
        self._optimizer.step(closure=closure)


class OptimizerWrapper(KeyedOptimizer):
    """Decorator base: forwards everything to the wrapped keyed optimizer and aliases its params / state / groups, so a subclass
    only overrides what it changes (``step`` for clipping, ``step`` + ``post_load_state_dict`` for warm-up ...)."""

    def __init__(self, optimizer: KeyedOptimizer) -> None:
        self._optimizer = optimizer
        self._alias()
        self.defaults: Dict[str, Any] = {"_save_param_groups": False}
        self._optimizer_step_pre_hooks: Dict[int, Callable] = {}
        self._optimizer_step_post_hooks: Dict[int, Callable] = {}

    def _alias(self) -> None:
        self.params: Mapping[str, ParamLike] = self._optimizer.params
        self.state: Mapping[Any, Any] = self._optimizer.state
        self.param_groups: Collection[Mapping[str, Any]] = self._optimizer.param_groups

    def __repr__(self) -> str:
        return repr(self._optimizer)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._optimizer.zero_grad(set_to_none=set_to_none)

    def step(self, closure: Any = None) -> None:
        self._optimizer.step(closure=closure)

    def add_param_group(self, param_group: Any) -> None:
        raise NotImplementedError("keyed optimizers are built over a fixed parameter map")

    def state_dict(self) -> Dict[str, Any]:
        return self._optimizer.state_dict()

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        self._optimizer.load_state_dict(state_dict)
        self.state = self._optimizer.state  # torch optimizers rebind these on load
        self.param_groups = self._optimizer.param_groups
        self.post_load_state_dict()

    def post_load_state_dict(self) -> None:
        self._optimizer.post_load_state_dict()

    def save_param_groups(self, save: bool) -> None:
        self._optimizer.save_param_groups(save)
