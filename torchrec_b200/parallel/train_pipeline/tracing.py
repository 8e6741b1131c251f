"""FX-based discovery of pipelineable modules and of where their inputs come from.

Reference: ``torchrec/distributed/train_pipeline/tracing.py`` - the ``*ArgInfoStep`` family :30-131, ``NodeArgsHelper`` :174-493, ``_get_leaf_module_names`` :496,
``Tracer`` :571-632; and ``utils.py: _rewrite_model`` :430-600.

The pipelines of this framework find sharded modules with forward pre-hooks on a first real batch (``train_pipelines.py: KJTGetter`` - no tracing, works
for models fx cannot trace). This module is the reference's static alternative: trace the model with sharded modules as leaves, walk every leaf call's
arguments back to the batch placeholder and express the walk as ``ArgInfo`` steps; anything between the batch and a sharded module that is itself a
module becomes a ``PipelinedPostproc`` candidate. ``rewrite_model`` returns the pipelined modules and their ``CallArgs`` - the same recipes the hook path
produces, so either can feed ``PipelinedForward``.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Set, Tuple, Type, Union

import torch
import torch.fx
from torch import nn

from ..types import ShardedModule
from .types import ArgInfo, BaseArgInfoStep, CallArgs, GetAttrArgInfoStep, GetItemArgInfoStep, PostprocArgInfoStep  # noqa: F401

logger = logging.getLogger(__name__)


class NoopArgInfoStep(BaseArgInfoStep):
    def process(self, arg: Any) -> Any:
        return arg


@dataclass
class ScalarArgInfoStep(BaseArgInfoStep):
    """A constant argument (python scalar / None) - ignores the batch."""

    value: Any

    def process(self, arg: Any) -> Any:
        return self.value


@dataclass
class ListArgInfoStep(BaseArgInfoStep):
    value: List[ArgInfo]

    def process(self, arg: Any) -> Any:
        return [v.process(arg) for v in self.value]


@dataclass
class DictArgInfoStep(BaseArgInfoStep):
    value: Dict[str, ArgInfo]

    def process(self, arg: Any) -> Any:
        return {k: v.process(arg) for k, v in self.value.items()}


class ArgInfoStepFactory:
    @classmethod
    def noop(cls) -> NoopArgInfoStep:
        return NoopArgInfoStep()

    @classmethod
    def get_attr(cls, name: str) -> GetAttrArgInfoStep:
        return GetAttrArgInfoStep(name)

    @classmethod
    def get_item(cls, index: Union[str, int]) -> GetItemArgInfoStep:
        return GetItemArgInfoStep(index)

    @classmethod
    def postproc(cls, module: nn.Module) -> PostprocArgInfoStep:
        return PostprocArgInfoStep(module)

    @classmethod
    def from_scalar(cls, value: Any) -> ScalarArgInfoStep:
        return ScalarArgInfoStep(value)

    @classmethod
    def from_list(cls, value: List[ArgInfo]) -> ListArgInfoStep:
        return ListArgInfoStep(value)

    @classmethod
    def from_dict(cls, value: Dict[str, ArgInfo]) -> DictArgInfoStep:
        return DictArgInfoStep(value)


def _get_leaf_module_names(model: nn.Module) -> List[str]:
    """FQNs the tracer must not look into - a SHALLOW trace, only as deep as pipelining needs: sharded modules, and the outermost modules that contain
    no sharded module (dense towers, postprocs) unless they opt in with ``_is_pytorch_fx_traceable = True``."""
    has_sharded: Dict[str, bool] = {}

    def visit(m: nn.Module, fqn: str) -> bool:
        found = isinstance(m, ShardedModule)
        for name, child in m.named_children():
            found = visit(child, f"{fqn}.{name}" if fqn else name) or found
        has_sharded[fqn] = found
        return found

    visit(model, "")
    leaves: List[str] = []
    for fqn, m in model.named_modules():
        if not fqn:
            continue
        parent = fqn.rsplit(".", 1)[0] if "." in fqn else ""
        if any(fqn.startswith(leaf + ".") for leaf in leaves):
            continue
        if isinstance(m, ShardedModule):
            leaves.append(fqn)
        elif not has_sharded[fqn] and not getattr(m, "_is_pytorch_fx_traceable", False) and has_sharded.get(parent, True):
            leaves.append(fqn)
    return leaves


class Tracer(torch.fx.Tracer):
    """fx tracer that treats sharded modules (and the given leaf FQNs) as opaque calls and never materialises parameters as proxies."""

    proxy_buffer_attributes = False

    def __init__(self, leaf_modules: Optional[List[str]] = None, extend_leaf_fqn: bool = False) -> None:
        super().__init__()
        self._leaf_modules: List[str] = list(leaf_modules or [])
        self._extend_leaf_fqn = extend_leaf_fqn

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        if isinstance(m, ShardedModule) or module_qualified_name in self._leaf_modules:
            return True
        if self._extend_leaf_fqn and any(module_qualified_name.startswith(leaf + ".") for leaf in self._leaf_modules):
            return True
        return super().is_leaf_module(m, module_qualified_name)


class NodeArgsHelper:
    """Turns the fx arguments of a leaf-module call into ``ArgInfo`` recipes rooted at the batch placeholder."""

    def __init__(self, model: nn.Module, pipelined_postprocs: Optional[Set[nn.Module]] = None, pipeline_postproc: bool = False,
                 postproc_wrapper: Optional[Callable[[nn.Module, str, List[ArgInfo], Dict[str, ArgInfo]], nn.Module]] = None) -> None:
        self._model = model
        self._pipeline_postproc = pipeline_postproc
        self._pipelined_postprocs: Set[nn.Module] = pipelined_postprocs if pipelined_postprocs is not None else set()
        self._postproc_wrapper = postproc_wrapper
        self._wrapped: Dict[str, nn.Module] = {}

    def _module(self, target: str) -> nn.Module:
        return self._model.get_submodule(target)

    def _walk(self, node: Any) -> Optional[List[BaseArgInfoStep]]:
        """Steps (batch -> value), or None when the value is not derivable from the batch by getattr / getitem / pipelineable postprocs."""
        if not isinstance(node, torch.fx.Node):
            return [ArgInfoStepFactory.from_scalar(node)] if isinstance(node, (int, float, bool, str, type(None))) else None
        if node.op == "placeholder":
            return []
        if node.op == "call_function" and node.target is getattr and len(node.args) == 2 and isinstance(node.args[1], str):
            up = self._walk(node.args[0])
            return None if up is None else up + [ArgInfoStepFactory.get_attr(node.args[1])]
        if node.op == "call_function" and getattr(node.target, "__name__", "") == "getitem" and len(node.args) == 2 and not isinstance(node.args[1], torch.fx.Node):
            up = self._walk(node.args[0])
            return None if up is None else up + [ArgInfoStepFactory.get_item(node.args[1])]
        if node.op == "call_method" and node.target in ("to", "contiguous", "detach") and node.args:
            return self._walk(node.args[0])  # device moves are the pipeline's job; value-preserving views pass through
        if node.op == "call_module" and self._pipeline_postproc:
            target = str(node.target)
            mod = self._module(target)
            if isinstance(mod, ShardedModule) or len(list(mod.parameters())) and any(p.requires_grad for p in mod.parameters()):
                return None  # trainable postprocs must run in the main forward (their weights change every step)
            if len(node.args) != 1 or node.kwargs:
                sub_args = [self.get_arg_info(a) for a in node.args]
                sub_kwargs = {k: self.get_arg_info(v) for k, v in node.kwargs.items()}
                if any(a is None for a in sub_args) or any(v is None for v in sub_kwargs.values()):
                    return None
                wrapped = self._wrap(mod, target, sub_args, sub_kwargs)  # type: ignore[arg-type]
                return [ArgInfoStepFactory.postproc(wrapped)]
            up = self._walk(node.args[0])
            if up is None:
                return None
            self._pipelined_postprocs.add(mod)
            return up + [ArgInfoStepFactory.postproc(mod)]
        return None

    def _wrap(self, mod: nn.Module, fqn: str, args: List[ArgInfo], kwargs: Dict[str, ArgInfo]) -> nn.Module:
        if fqn not in self._wrapped:
            self._wrapped[fqn] = self._postproc_wrapper(mod, fqn, args, kwargs) if self._postproc_wrapper is not None else _CallWithRecipes(mod, args, kwargs)
            self._pipelined_postprocs.add(mod)
        return self._wrapped[fqn]

    def get_arg_info(self, arg: Any) -> Optional[ArgInfo]:
        if isinstance(arg, (list, tuple)) and any(isinstance(a, torch.fx.Node) for a in arg):
            items = [self.get_arg_info(a) for a in arg]
            return None if any(i is None for i in items) else ArgInfo([ArgInfoStepFactory.from_list(items)])  # type: ignore[arg-type]
        if isinstance(arg, dict) and any(isinstance(a, torch.fx.Node) for a in arg.values()):
            items_d = {k: self.get_arg_info(v) for k, v in arg.items()}
            return None if any(i is None for i in items_d.values()) else ArgInfo([ArgInfoStepFactory.from_dict(items_d)])  # type: ignore[arg-type]
        steps = self._walk(arg)
        return None if steps is None else ArgInfo(steps)

    def get_node_args(self, node: torch.fx.Node) -> Tuple[CallArgs, int]:
        """(recipes of all arguments that could be resolved, number resolved). A call is pipelineable iff all of its arguments resolved."""
        args, kwargs, found = [], {}, 0
        for a in node.args:
            info = self.get_arg_info(a)
            if info is not None:
                args.append(info)
                found += 1
        for k, v in node.kwargs.items():
            info = self.get_arg_info(v)
            if info is not None:
                kwargs[k] = info
                found += 1
        return CallArgs(args=args, kwargs=kwargs), found


class _CallWithRecipes(nn.Module):
    """A multi-input postproc applied to the batch: builds its own arguments from recipes."""

    def __init__(self, mod: nn.Module, args: List[ArgInfo], kwargs: Dict[str, ArgInfo]) -> None:
        super().__init__()
        self._mod, self._args, self._kwargs = mod, args, kwargs

    def forward(self, batch: Any) -> Any:
        return self._mod(*[a.process(batch) for a in self._args], **{k: v.process(batch) for k, v in self._kwargs.items()})


@dataclass
class PipelinedModuleInfo:
    fqn: str
    module: nn.Module
    call_args: CallArgs


def rewrite_model(model: nn.Module, batch: Any = None, pipeline_postproc: bool = False, leaf_modules: Optional[List[str]] = None,
                  concrete_args: Optional[Dict[str, Any]] = None) -> Tuple[List[PipelinedModuleInfo], List[str], Set[nn.Module]]:
    """Trace ``model`` and return (pipelineable sharded modules with their input recipes, FQNs of sharded modules that are NOT pipelineable, the
    postproc modules folded into recipes). Unwraps ``DistributedModelParallel`` / DDP like the reference's ``_rewrite_model``."""
    inner = model
    while hasattr(inner, "_dmp_wrapped_module") or (hasattr(inner, "module") and isinstance(getattr(inner, "module"), nn.Module) and type(inner).__name__ in (
            "DistributedModelParallel", "DistributedDataParallel", "DMPCollection")):
        inner = getattr(inner, "_dmp_wrapped_module", None) or inner.module
    sharded = {fqn: m for fqn, m in inner.named_modules() if isinstance(m, ShardedModule)}
    if not sharded:
        return [], [], set()
    tracer = Tracer(leaf_modules=(leaf_modules or []) + _get_leaf_module_names(inner))
    graph = tracer.trace(inner, concrete_args=concrete_args)
    postprocs: Set[nn.Module] = set()
    helper = NodeArgsHelper(inner, postprocs, pipeline_postproc)
    ok: List[PipelinedModuleInfo] = []
    bad: List[str] = []
    for node in graph.nodes:
        if node.op == "call_module" and str(node.target) in sharded:
            call_args, found = helper.get_node_args(node)
            total = len(node.args) + len(node.kwargs)
            if found == total and total > 0:
                ok.append(PipelinedModuleInfo(str(node.target), sharded[str(node.target)], call_args))
            else:
                bad.append(str(node.target))
                logger.warning("module %s cannot be pipelined: %d of %d inputs are not derivable from the batch", node.target, total - found, total)
    return ok, bad, postprocs


_rewrite_model = rewrite_model
