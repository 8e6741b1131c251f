"""Helpers around quantized modules for tracing / packaging (reference torchrec/quant/utils.py:23-116)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .embedding_modules import EmbeddingBagCollection as QuantEmbeddingBagCollection
from .embedding_modules import EmbeddingCollection as QuantEmbeddingCollection


def populate_fx_names(quant_ebc: nn.Module) -> None:
    """Give the kernel modules of a quantized (sharded or not) collection a stable ``_fx_path``: a tracer that meets an unregistered
    table-batched module falls back to that name (``emb_module.<tables>`` unsharded, ``embedding_lookup.rank_<r>.tbe`` sharded)."""
    tbes = getattr(quant_ebc, "_tbes", None)
    if tbes is None and hasattr(quant_ebc, "_tbe"):
        tbes = [quant_ebc._tbe]
    if tbes is None:
        return
    sharded = type(quant_ebc).__name__.startswith("Sharded")
    for i, tbe in enumerate(tbes):
        if not hasattr(tbe, "embedding_specs"):
            continue
        if sharded:
            tbe._fx_path = f"embedding_lookup.rank_{i}.tbe"
        else:
            tbe._fx_path = "emb_module." + ",".join(str(s[0]) for s in tbe.embedding_specs)


def recursive_populate_fx_names(module: nn.Module) -> None:
    if hasattr(module, "_tbes") or (hasattr(module, "_tbe") and isinstance(module, (QuantEmbeddingCollection,))):
        populate_fx_names(module)
        return
    for sub in module.children():
        recursive_populate_fx_names(sub)


def meta_to_cpu_placement(module: nn.Module) -> None:
    """Replace quantized collections that still live on the meta device by CPU instances of the same configuration (weights are filled
    afterwards from a state dict): lets a model planned on ``meta`` be packaged / traced on a host without GPUs."""
    if hasattr(module, "_dmp_wrapped_module"):
        _meta_to_cpu_placement(module.module, module, "_dmp_wrapped_module")
    else:
        _meta_to_cpu_placement(module, module)


def _meta_to_cpu_placement(module: nn.Module, root_module: nn.Module, name: Optional[str] = None) -> None:
    if name is not None and isinstance(module, QuantEmbeddingBagCollection) and module.device.type == "meta":
        setattr(root_module, name, QuantEmbeddingBagCollection(tables=module.embedding_bag_configs(), is_weighted=module.is_weighted(), device=torch.device("cpu"),
                                                               output_dtype=module.output_dtype(), row_alignment=module.row_alignment))
    elif name is not None and isinstance(module, QuantEmbeddingCollection) and module.device.type == "meta":
        setattr(root_module, name, QuantEmbeddingCollection(tables=module.embedding_configs(), device=torch.device("cpu"), need_indices=module.need_indices(),
                                                            output_dtype=module.output_dtype(), row_alignment=getattr(module, "row_alignment", 16)))
    else:
        for child_name, sub in module.named_children():
            _meta_to_cpu_placement(sub, module, child_name)
