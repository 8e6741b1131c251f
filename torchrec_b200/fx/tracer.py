"""FX tracing with recsys-aware leaves (reference torchrec/fx/tracer.py:30-180, fx/utils.py).
Sharded modules, embedding collections and feature processors stay opaque ``call_module`` nodes; KJT / KT method calls on
proxies are recorded as ``call_method`` nodes. The train pipelines do NOT need tracing (they match sharded modules by
their KJT argument), the tracer exists for model surgery and for exporting dense sub-graphs."""
from __future__ import annotations

import threading
from typing import Any, Callable, Dict, List, Optional, Union

import torch
import torch.fx
from torch import nn

_state = threading.local()


def is_fx_tracing() -> bool:
    return getattr(_state, "tracing", False) or torch.fx._symbolic_trace.is_fx_tracing()


class Tracer(torch.fx.Tracer):
    def __init__(self, leaf_modules: Optional[List[str]] = None) -> None:
        super().__init__()
        self._leaf_modules: List[str] = leaf_modules if leaf_modules is not None else []

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        from ..modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
        from ..modules.feature_processor_ import FeatureProcessor, FeatureProcessorsCollection
        from ..parallel.types import ShardedModule

        if isinstance(m, (ShardedModule, EmbeddingBagCollection, EmbeddingCollection, FeatureProcessor, FeatureProcessorsCollection)):
            return True
        if type(m).__name__ in self._leaf_modules or module_qualified_name in self._leaf_modules:
            return True
        return super().is_leaf_module(m, module_qualified_name)

    def trace(self, root: Union[nn.Module, Callable[..., Any]], concrete_args: Optional[Dict[str, Any]] = None) -> torch.fx.Graph:
        _state.tracing = True
        try:
            return super().trace(root, concrete_args)
        finally:
            _state.tracing = False

    # KJT-family constants and already-resolved awaitables that reach ``create_arg`` (a model closing over a fixed KJT, a NoWait wrapped
    # around a proxy) are not fx base types: constants become ``get_attr`` nodes on the root (so the GraphModule carries them),
    # ``NoWait(x)`` traces as ``x`` (reference fx/tracer.py:95-128)
    def create_arg(self, a: Any) -> Any:
        from ..parallel.types import NoWait
        from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor

        if isinstance(a, NoWait):
            return self.create_arg(a._obj)
        if isinstance(a, (KeyedJaggedTensor, JaggedTensor, KeyedTensor)):
            root = getattr(self, "root", None)
            if isinstance(root, nn.Module):
                n = 0
                while hasattr(root, f"_sparse_constant_{n}"):
                    if getattr(root, f"_sparse_constant_{n}") is a:
                        break
                    n += 1
                name = f"_sparse_constant_{n}"
                if not hasattr(root, name):
                    object.__setattr__(root, name, a)
                return self.create_node("get_attr", name, (), {})
        if isinstance(a, torch.device):  # (plain fx accepts devices only through repr-able constants)
            return a
        return super().create_arg(a)

    def path_of_module(self, mod: nn.Module) -> str:
        try:
            return super().path_of_module(mod)
        except NameError:
            # modules created on the fly inside forward (e.g. KeyedTensor helpers)
            return f"_dynamic_{type(mod).__name__}"


def symbolic_trace(root: Union[nn.Module, Callable], concrete_args: Optional[Dict[str, Any]] = None, leaf_modules: Optional[List[str]] = None) -> torch.fx.GraphModule:
    tracer = Tracer(leaf_modules)
    graph = tracer.trace(root, concrete_args)
    return torch.fx.GraphModule(root if isinstance(root, nn.Module) else nn.Module(), graph)


from .utils import assert_fx_safe, dmp_fx_trace_forward, fake_range, fx_marker, is_marker_node  # noqa: E402,F401
