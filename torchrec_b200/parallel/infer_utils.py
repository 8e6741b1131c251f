"""Introspection of sharded / quantized models for serving (reference torchrec/distributed/infer_utils.py:30-196)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Set, Tuple, Type

import torch
from torch import nn


def _is_tbe(m: nn.Module) -> bool:
    return hasattr(m, "embedding_specs") and hasattr(m, "feature_table_map")


def get_tbes_from_sharded_module(module: nn.Module) -> List[nn.Module]:
    """Every table-batched kernel module under ``module`` (training TBEs, cached wrappers' backing tables, quantized inference TBEs)."""
    return [m for m in module.modules() if _is_tbe(m)]


def get_tbe_specs_from_sharded_module(module: nn.Module) -> List[Tuple[str, int, int, str, str]]:
    """``(table name, rows, cols, weight dtype, location)`` of every local table shard."""
    out: List[Tuple[str, int, int, str, str]] = []
    for tbe in get_tbes_from_sharded_module(module):
        names = getattr(tbe, "table_names", None) or [f"t{i}" for i in range(len(tbe.embedding_specs))]
        w = getattr(tbe, "weights", None)
        dtype = str(getattr(tbe, "weights_precision", None) or (w.dtype if isinstance(w, torch.Tensor) else "unknown")).replace("torch.", "")
        loc = getattr(getattr(tbe, "location", None), "name", "DEVICE")
        for n, spec in zip(names, tbe.embedding_specs):
            rows, cols = int(spec[0]), int(spec[1])
            out.append((str(n), rows, cols, dtype, loc))
    return out


def get_path_device_tuples(module: nn.Module, ignore_list: Optional[List[str]] = None) -> List[Tuple[str, str]]:
    """``(parameter / buffer path, device)`` for everything in ``module``: a quick check that a sharded inference model placed its
    shards on the devices the plan names."""
    ignore = set(ignore_list or [])
    out: List[Tuple[str, str]] = []
    for name, t in list(module.named_parameters(remove_duplicate=False)) + list(module.named_buffers(remove_duplicate=False)):
        if any(tok in name for tok in ignore):
            continue
        out.append((name, str(t.device)))
    return sorted(set(out))


def get_all_torchrec_modules(model: nn.Module, trec_module_class_types: Optional[List[Type[nn.Module]]] = None) -> Dict[str, nn.Module]:
    """fqn -> module for every embedding collection / sharded module in ``model``."""
    from ..modules.embedding_modules import EmbeddingBagCollectionInterface, EmbeddingCollectionInterface
    from .types import ShardedModule

    kinds: Tuple[type, ...] = tuple(trec_module_class_types or [EmbeddingBagCollectionInterface, EmbeddingCollectionInterface, ShardedModule])
    return {name: m for name, m in model.named_modules() if isinstance(m, kinds)}


def get_non_scriptable_trec_module(model: nn.Module) -> Dict[str, nn.Module]:
    """Sharded modules are driven by the Python engine (ctypes kernels, NVLink buffers): none of them is TorchScript-able; they are the
    leaves an exporter has to treat as opaque."""
    from .types import ShardedModule

    return {name: m for name, m in model.named_modules() if isinstance(m, ShardedModule)}
