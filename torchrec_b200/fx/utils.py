"""Helpers around fx tracing of recommender models (reference ``torchrec/fx/utils.py``): tracing the module behind a
``DistributedModelParallel`` with its sharded leaves intact, marker nodes that survive graph transformations, trace-safe assertions."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
import torch.fx
from torch import nn


def fake_range() -> List[int]:
    """A loop range that does not depend on a traced value (``for _ in fake_range()`` traces exactly one iteration)."""
    return [0]


torch.fx.wrap("fake_range")


def dmp_fx_trace_forward(dmp: nn.Module, tracer: Optional[torch.fx.Tracer] = None) -> torch.fx.GraphModule:
    """fx GraphModule of the model wrapped by a ``DistributedModelParallel``: the DDP / FSDP wrapper is stripped first, sharded
    modules stay ``call_module`` leaves (they own streams, collectives and awaitables that must not be traced into), so the graph is
    the dense skeleton + opaque sparse calls - what pipeline rewriting and dense-subgraph export operate on."""
    from ..parallel.model_parallel import get_unwrapped_module
    from .tracer import Tracer

    module = get_unwrapped_module(dmp)
    tracer = tracer if tracer is not None else Tracer()
    graph = tracer.trace(module)
    gm = torch.fx.GraphModule(module, graph)
    # the sharded leaves of the traced module are shared with the live model (same objects): calling ``gm`` trains the same tables
    return gm


def _fx_marker(s: str, any_proxy_unused: Any) -> None:
    return None


torch.fx.wrap("_fx_marker")


def fx_marker(s: str, any_proxy_unused: Any) -> None:
    """Leave a named no-op node in the traced graph (``fx_marker("KJT_ONE_TO_ALL_FORWARD_BEGIN", kjt)``): graph passes find regions
    by name (``is_marker_node``) instead of by structure. ``any_proxy_unused`` ties the node to the data flow so dead-code
    elimination ordering keeps it between its neighbours."""
    _fx_marker(s, any_proxy_unused)


def is_marker_node(node: torch.fx.Node, marker_name: str) -> bool:
    return bool(node.op == "call_function" and getattr(node.target, "__name__", "") == "_fx_marker" and len(node.args) > 0 and node.args[0] == marker_name)


def marker_regions(graph: torch.fx.Graph, begin: str, end: str) -> List[List[torch.fx.Node]]:
    """Node lists between every ``begin`` / ``end`` marker pair, in graph order."""
    regions: List[List[torch.fx.Node]] = []
    cur: Optional[List[torch.fx.Node]] = None
    for n in graph.nodes:
        if is_marker_node(n, begin):
            cur = []
        elif is_marker_node(n, end):
            if cur is not None:
                regions.append(cur)
            cur = None
        elif cur is not None:
            cur.append(n)
    return regions


def assert_fx_safe(condition: Any, message: str) -> None:
    """``assert`` that is skipped while tracing (a proxy has no truth value; data-dependent checks only make sense eagerly)."""
    from .tracer import is_fx_tracing

    if not is_fx_tracing() and not isinstance(condition, torch.fx.Proxy):
        assert condition, message


def leaf_call_counts(gm: torch.fx.GraphModule) -> Dict[str, int]:
    """How often each leaf module is called in a traced graph (a sharded module called twice needs two pipeline contexts)."""
    counts: Dict[str, int] = {}
    for n in gm.graph.nodes:
        if n.op == "call_module":
            counts[str(n.target)] = counts.get(str(n.target), 0) + 1
    return counts
