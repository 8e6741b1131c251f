"""Export-time helpers of the IR package: opaque placeholder ops for sparse modules, dynamic-shape marking of KJT inputs and graph
clean-ups of unflattened programs.

``torch.export`` cannot (and should not) trace through an embedding collection: the sparse modules are *encapsulated* - their
configuration is serialized next to them (``serializer.encapsulate_ir_modules``) and, while the program is captured, their forward is
swapped for a placeholder op that only promises output shapes (``torchrec_b200::ir_emb_lookup`` etc.). After loading, the placeholder
modules are rebuilt from the serialized configs (``decapsulate_ir_modules``) - sharded, quantized or plain - and run the real kernels.

Parity: reference ``torchrec/ir/utils.py`` (custom ops :54-133, ``encapsulate_ir_modules`` :135, ``decapsulate_ir_modules`` :166,
``mark_dynamic_kjt`` :216, ``move_to_copy_nodes_to_device`` :298, ``prune_pytree_flatten_unflatten`` :374)."""
from __future__ import annotations

import operator
from typing import Any, Callable, Dict, List, Optional, Tuple, Type, Union

import torch
from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor
from .serializer import JsonSerializer, decapsulate_ir_modules, encapsulate_ir_modules, serialize_sparse_modules  # noqa: F401


def qualname(m: Union[nn.Module, Type[nn.Module]]) -> str:
    cls = type(m) if isinstance(m, nn.Module) else m
    return f"{cls.__module__}.{cls.__qualname__}"


def get_device(tensors: List[Optional[torch.Tensor]]) -> Optional[torch.device]:
    """Device of the first real tensor of a flattened input list (None when there is none)."""
    for t in tensors:
        if t is not None:
            return t.device
    return None


# ---- placeholder ops ------------------------------------------------------------------------------------------------------------
# One output per entry of ``dims``: [batch_size, dim]. The eager implementations return zeros: an exported program that still contains
# them was never decapsulated, and zeros make that obvious without crashing shape propagation.
def _placeholder(tensors: List[Optional[torch.Tensor]], batch_size: int, dims: List[int]) -> List[torch.Tensor]:
    dev = get_device(tensors)
    return [torch.zeros(batch_size, d, device=dev) for d in dims]


@torch.library.custom_op("torchrec_b200::ir_emb_lookup", mutates_args=())
def ir_emb_lookup(tensors: List[Optional[torch.Tensor]], batch_size: int, dims: List[int]) -> List[torch.Tensor]:
    return _placeholder(tensors, batch_size, dims)


@ir_emb_lookup.register_fake
def _(tensors, batch_size, dims):
    dev = get_device(tensors)
    return [torch.empty(batch_size, d, device=dev) for d in dims]


@torch.library.custom_op("torchrec_b200::ir_kt_regroup", mutates_args=())
def ir_kt_regroup(tensors: List[Optional[torch.Tensor]], batch_size: int, dims: List[int]) -> List[torch.Tensor]:
    return _placeholder(tensors, batch_size, dims)


@ir_kt_regroup.register_fake
def _(tensors, batch_size, dims):
    dev = get_device(tensors)
    return [torch.empty(batch_size, d, device=dev) for d in dims]


@torch.library.custom_op("torchrec_b200::ir_dynamic_batch_emb_lookup", mutates_args=())
def ir_dynamic_batch_emb_lookup(tensors: List[Optional[torch.Tensor]], batch_size: torch.Tensor, dims: List[int]) -> List[torch.Tensor]:
    """Variable batch: the batch size is carried by the first dimension of ``batch_size`` (a data-dependent size stays symbolic)."""
    return _placeholder(tensors, batch_size.shape[0], dims)


@ir_dynamic_batch_emb_lookup.register_fake
def _(tensors, batch_size, dims):
    dev = get_device(tensors)
    return [torch.empty(batch_size.shape[0], d, device=dev) for d in dims]


def _kjt_tensors(kjt: KeyedJaggedTensor) -> List[Optional[torch.Tensor]]:
    return [kjt.values(), kjt.lengths_or_none(), kjt.offsets_or_none(), kjt.weights_or_none()]


def ebc_placeholder_forward(module: nn.Module) -> Callable[[KeyedJaggedTensor], KeyedTensor]:
    """Export-time forward of an (encapsulated) EmbeddingBagCollection: one placeholder op, a KeyedTensor of the right keys / widths."""
    cfgs = module.embedding_bag_configs()
    keys = [n for c in cfgs for n in (getattr(c, "embedding_names", None) or c.feature_names)]
    dims = [c.embedding_dim for c in cfgs for _ in c.feature_names]

    def forward(features: KeyedJaggedTensor) -> KeyedTensor:
        outs = torch.ops.torchrec_b200.ir_emb_lookup(_kjt_tensors(features), features.stride(), [sum(dims)])
        return KeyedTensor(keys=keys, length_per_key=dims, values=outs[0])

    return forward


def swap_placeholder_forwards(model: nn.Module) -> Dict[str, Callable[..., Any]]:
    """Every module carrying ``ir_metadata`` whose type has a placeholder forward gets it installed; returns the originals by fqn
    (``restore_forwards`` puts them back)."""
    originals: Dict[str, Callable[..., Any]] = {}
    for fqn, m in model.named_modules():
        if getattr(m, "ir_metadata", None) is None:
            continue
        if hasattr(m, "embedding_bag_configs") and not hasattr(m, "_feature_processors"):
            originals[fqn] = m.forward
            m.forward = ebc_placeholder_forward(m)  # type: ignore[method-assign]
    return originals


def restore_forwards(model: nn.Module, originals: Dict[str, Callable[..., Any]]) -> None:
    for fqn, fwd in originals.items():
        model.get_submodule(fqn).forward = fwd  # type: ignore[method-assign]


# ---- dynamic shapes of KJT inputs -----------------------------------------------------------------------------------------------------
def _dim(name: str, min: Optional[int] = None, max: Optional[int] = None):
    from torch.export import Dim

    kw: Dict[str, int] = {}
    if min is not None:
        kw["min"] = min
    if max is not None:
        kw["max"] = max
    return Dim(name, **kw)


_DIM_COUNTER = {"n": 0}


def mark_dynamic_kjt(kjt: KeyedJaggedTensor, shapes_collection=None, variable_length: bool = False, vlen=None, llen=None, variable_batch: bool = False):
    """Register the data-dependent dimensions of a KJT input with a ``torch.export.ShapesCollection``: the number of values (shared by
    ``values`` and ``weights``) always, the number of bags (``lengths`` / ``offsets``) when ``variable_length`` (variable batch). The
    collection is what ``torch.export.export(..., dynamic_shapes=collection)`` takes. ``vlen`` / ``llen``: reuse Dims across KJTs that
    must agree."""
    from torch.export import ShapesCollection

    variable_length = variable_length or variable_batch  # the reference's name of the flag
    if shapes_collection is None:
        shapes_collection = ShapesCollection()
    _DIM_COUNTER["n"] += 1
    n = _DIM_COUNTER["n"]
    vlen = vlen if vlen is not None else _dim(f"vlen{n}", min=2)
    if kjt.values().numel() >= 2:
        shapes_collection[kjt.values()] = (vlen,)
        w = kjt.weights_or_none()
        if w is not None and w.numel() >= 2:
            shapes_collection[w] = (vlen,)
    if variable_length:
        llen = llen if llen is not None else _dim(f"llen{n}", min=2)
        lengths, offsets = kjt.lengths_or_none(), kjt.offsets_or_none()
        if lengths is not None and lengths.numel() >= 2:
            shapes_collection[lengths] = (llen,)
        if offsets is not None and offsets.numel() >= 3:
            shapes_collection[offsets] = (llen + 1,)
    return shapes_collection


# ---- graph clean-ups of unflattened programs -----------------------------------------------------------------------------------------------
def move_to_copy_nodes_to_device(unflattened_module: nn.Module, device: torch.device) -> nn.Module:
    """An exported program bakes the capture-time device into its ``aten._to_copy`` nodes; re-target them (e.g. capture on cpu / meta,
    serve on cuda:k)."""
    for m in unflattened_module.modules():
        graph = getattr(m, "graph", None)
        if graph is None:
            continue
        changed = False
        for node in graph.nodes:
            if node.op == "call_function" and "_to_copy" in str(node.target) and "device" in node.kwargs:
                kw = dict(node.kwargs)
                kw["device"] = device
                node.kwargs = kw
                changed = True
        if changed and hasattr(m, "recompile"):
            m.recompile()
    return unflattened_module


def _is_flatten(node) -> bool:
    return node.op == "call_function" and getattr(node.target, "__name__", "") in ("tree_flatten_spec", "fx_pytree_tree_flatten_spec")


def _is_unflatten(node) -> bool:
    return node.op == "call_function" and getattr(node.target, "__name__", "") in ("tree_unflatten",)


def prune_pytree_flatten_unflatten(module: nn.Module, in_place: bool = False) -> nn.Module:
    """``unflatten`` leaves ``tree_flatten_spec`` at the top of every sub-graph and ``tree_unflatten`` at its end. When a sub-module's
    inputs arrive already flat (positional tensors) the pair is an identity: users of ``flatten(args)[i]`` are rewired to the i-th
    placeholder and a trailing ``unflatten`` of a single-leaf spec to its only leaf. Graphs where that does not hold are left alone."""
    import copy

    mod = module if in_place else copy.deepcopy(module)
    for m in mod.modules():
        graph = getattr(m, "graph", None)
        if graph is None:
            continue
        placeholders = [n for n in graph.nodes if n.op == "placeholder"]
        changed = False
        for node in list(graph.nodes):
            if _is_flatten(node):
                users = list(node.users)
                if not users or not all(u.op == "call_function" and u.target is operator.getitem and isinstance(u.args[1], int) and u.args[1] < len(placeholders)
                                        for u in users):
                    continue
                for u in users:
                    u.replace_all_uses_with(placeholders[u.args[1]])
                    graph.erase_node(u)
                graph.erase_node(node)
                changed = True
            elif _is_unflatten(node):
                leaves = node.args[0] if node.args else None
                if isinstance(leaves, (list, tuple)) and len(leaves) == 1:
                    node.replace_all_uses_with(leaves[0])
                    graph.erase_node(node)
                    changed = True
        if changed:
            graph.eliminate_dead_code()
            if hasattr(m, "recompile"):
                m.recompile()
    return mod
