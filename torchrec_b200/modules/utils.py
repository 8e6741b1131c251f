"""Small helpers shared by modules (reference torchrec/modules/utils.py)."""
from __future__ import annotations

import copy
from collections import defaultdict
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Tuple, Union

import torch

from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor


def get_module_output_dimension(module: Union[Callable, torch.nn.Module], in_features: int) -> int:
    inp = torch.zeros(1, in_features)
    return module(inp).size(-1)


def check_module_output_dimension(module, in_features: int, out_features: int) -> bool:
    if isinstance(module, (list, torch.nn.ModuleList)):
        return all(check_module_output_dimension(m, in_features, out_features) for m in module)
    return get_module_output_dimension(module, in_features) == out_features


def extract_module_or_tensor_callable(module_or_callable):
    try:
        module = module_or_callable()
        if isinstance(module, torch.nn.Module):
            return module
        raise ValueError("Expected callable that takes no input to return a torch.nn.Module")
    except TypeError as e:
        if "required positional argument" in str(e):
            return module_or_callable
        raise


def convert_list_of_modules_to_modulelist(modules: Iterable[torch.nn.Module], sizes: Tuple[int, ...]) -> torch.nn.Module:
    assert len(sizes) == 1 or len(sizes) == 2
    mods = list(modules)
    assert len(mods) == sizes[0]
    if len(sizes) == 1:
        return torch.nn.ModuleList(mods)
    for m in mods:
        assert len(m) == sizes[1]  # type: ignore[arg-type]
    return torch.nn.ModuleList(torch.nn.ModuleList(m) for m in mods)


def construct_jagged_tensors(
    embeddings: torch.Tensor,
    features: KeyedJaggedTensor,
    embedding_names: List[str],
    need_indices: bool = False,
    features_to_permute_indices: Optional[Dict[str, List[int]]] = None,
    original_features: Optional[KeyedJaggedTensor] = None,
    reverse_indices: Optional[torch.Tensor] = None,
    seq_vbe_ctx=None,
) -> Dict[str, JaggedTensor]:
    """Split a [sum L, D] sequence-embedding tensor back into per-feature JaggedTensors."""
    if original_features is not None:
        features = original_features
    if reverse_indices is not None:
        embeddings = torch.index_select(embeddings, 0, reverse_indices.to(torch.int32))
    ret: Dict[str, JaggedTensor] = {}
    stride = features.stride()
    length_per_key = features.length_per_key()
    values = features.values()
    lengths = features.lengths().view(-1, stride) if not features.variable_stride_per_key() else None
    lo = features.lengths_offset_per_key()
    embeddings_list = torch.split(embeddings, length_per_key, dim=0)
    values_list = torch.split(values, length_per_key) if need_indices else None
    key_indices = defaultdict(list)
    for i, key in enumerate(embedding_names):
        key_indices[key].append(i)
    for i, key in enumerate(features.keys()):
        if key not in key_indices:
            continue
        ls = lengths[i] if lengths is not None else features.lengths()[lo[i] : lo[i + 1]]
        ret[key] = JaggedTensor(lengths=ls, values=embeddings_list[i], weights=values_list[i] if need_indices else None)
    return ret


@dataclass
class SequenceVBEContext:
    recat: torch.Tensor
    unpadded_lengths: torch.Tensor
    reindexed_lengths: torch.Tensor
    reindexed_length_per_key: List[int]
    reindexed_values: Optional[torch.Tensor] = None


def deterministic_dedup(ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    sorted_id_values, sorted_id_indices = ids.sort()
    sorted_unique_ids, sorted_unique_inverses = sorted_id_values.unique_consecutive(return_counts=False, return_inverse=True)
    last_existence_index = torch.scatter_reduce(
        torch.full_like(sorted_unique_ids, -1), 0, sorted_unique_inverses, sorted_id_indices, "amax")
    return sorted_unique_ids.view(-1), last_existence_index.flatten()


def jagged_index_select_with_empty(values, ids, offsets, output_offsets) -> torch.Tensor:
    if ids.size(0) == 0:
        return torch.empty(0, device=values.device, dtype=values.dtype)
    from ..ops.jagged import jagged_index_select_2d

    lengths = offsets[1:] - offsets[:-1]
    out, _ = jagged_index_select_2d(values.flatten().unsqueeze(-1), lengths, ids)
    return out.flatten()
