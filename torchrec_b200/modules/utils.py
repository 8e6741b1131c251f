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


def _permute_indices(indices: List[int], permute: List[int]) -> List[int]:
    out = [0] * len(indices)
    for i, src in enumerate(permute):
        out[i] = indices[src]
    return out


def _vbe_reindex(embeddings: torch.Tensor, seq_vbe_ctx: "SequenceVBEContext"):
    """Variable batch per feature: re-order the looked-up rows segment by segment (``recat`` over the un-padded lengths) into the
    (feature, sample) order of the re-indexed lengths."""
    from ..ops import jagged as J

    dim = embeddings.shape[1]
    total = sum(seq_vbe_ctx.reindexed_length_per_key)
    _, rows, _ = J.permute_1D_sparse_data(seq_vbe_ctx.recat, seq_vbe_ctx.unpadded_lengths.reshape(-1), embeddings, None, total)
    assert seq_vbe_ctx.reindexed_lengths.dim() == 2
    return rows.view(-1, dim), seq_vbe_ctx.reindexed_lengths, seq_vbe_ctx.reindexed_length_per_key, seq_vbe_ctx.reindexed_values


def _assemble_jt_dict(embeddings, lengths, values, length_per_key, embedding_names, need_indices, features_to_permute_indices) -> Dict[str, JaggedTensor]:
    """One JaggedTensor per distinct embedding name; a name that occurs several times is a column-wise sharded table whose column
    blocks (re-ordered by ``features_to_permute_indices``) are concatenated along the embedding dimension."""
    lengths_tuple = torch.unbind(lengths, dim=0)
    embeddings_list = torch.split(embeddings, length_per_key, dim=0)
    values_list = torch.split(values, length_per_key) if need_indices and values is not None else None
    key_indices: Dict[str, List[int]] = defaultdict(list)
    for i, key in enumerate(embedding_names):
        key_indices[key].append(i)
    ret: Dict[str, JaggedTensor] = {}
    for key, indices in key_indices.items():
        if features_to_permute_indices and key in features_to_permute_indices:
            indices = _permute_indices(indices, features_to_permute_indices[key])
        ret[key] = JaggedTensor(lengths=lengths_tuple[indices[0]],
                                values=embeddings_list[indices[0]] if len(indices) == 1 else torch.cat([embeddings_list[i] for i in indices], dim=1),
                                weights=values_list[indices[0]] if values_list is not None else None)
    return ret


def construct_jagged_tensors(
    embeddings: torch.Tensor,
    features: KeyedJaggedTensor,
    embedding_names: List[str],
    need_indices: bool = False,
    features_to_permute_indices: Optional[Dict[str, List[int]]] = None,
    original_features: Optional[KeyedJaggedTensor] = None,
    reverse_indices: Optional[torch.Tensor] = None,
    seq_vbe_ctx=None,
    use_gather_select: bool = False,
) -> Dict[str, JaggedTensor]:
    """Split a [sum L, D] sequence-embedding tensor back into per-feature JaggedTensors (reference modules/utils.py:246).
    ``embedding_names[i]`` names the i-th key of ``features`` (a column-wise sharded table occurs once per column block);
    ``reverse_indices`` undoes an id de-duplication; ``need_indices`` puts the ids into the weights slot."""
    if original_features is not None:
        features = original_features
    if reverse_indices is not None:
        if use_gather_select:  # gather has the cheaper backward in many cases
            embeddings = torch.gather(embeddings, 0, reverse_indices.long().unsqueeze(1).expand(-1, embeddings.size(-1)))
        else:
            embeddings = torch.index_select(embeddings, 0, reverse_indices.to(torch.int32) if reverse_indices.dtype != torch.int64 else reverse_indices)
    if seq_vbe_ctx is not None:
        embeddings, lengths, length_per_key, values = _vbe_reindex(embeddings, seq_vbe_ctx)
    elif features.variable_stride_per_key():
        # variable batch without a re-index context: one row of lengths per key of its own size
        lo = features.lengths_offset_per_key()
        flat = features.lengths()
        ret: Dict[str, JaggedTensor] = {}
        embeddings_list = torch.split(embeddings, features.length_per_key(), dim=0)
        values_list = torch.split(features.values(), features.length_per_key()) if need_indices else None
        for i, key in enumerate(embedding_names):
            ret[key] = JaggedTensor(lengths=flat[lo[i] : lo[i + 1]], values=embeddings_list[i], weights=values_list[i] if values_list is not None else None)
        return ret
    else:
        lengths = features.lengths().view(-1, features.stride())
        length_per_key = features.length_per_key()
        values = features.values()
    return _assemble_jt_dict(embeddings, lengths, values, length_per_key, embedding_names, need_indices, features_to_permute_indices)


def construct_jagged_tensors_inference(
    embeddings: torch.Tensor,
    lengths: torch.Tensor,
    values: torch.Tensor,
    embedding_names: List[str],
    need_indices: bool = False,
    features_to_permute_indices: Optional[Dict[str, List[int]]] = None,
    reverse_indices: Optional[torch.Tensor] = None,
    remove_padding: bool = False,
) -> Dict[str, JaggedTensor]:
    """The inference form: ``lengths`` [F, B] and ``values`` are passed as tensors (no KJT in the traced graph); ``remove_padding`` cuts
    rows that a fixed-size lookup appended after the real ones."""
    if reverse_indices is not None:
        embeddings = torch.index_select(embeddings, 0, reverse_indices.long())
    elif remove_padding:
        embeddings = embeddings[: int(lengths.sum())]
    length_per_key = torch.sum(lengths, dim=1).tolist()
    return _assemble_jt_dict(embeddings, lengths, values, length_per_key, embedding_names, need_indices, features_to_permute_indices)


def init_mlp_weights_xavier_uniform(m: torch.nn.Module) -> None:
    if isinstance(m, torch.nn.Linear):
        torch.nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            m.bias.data.fill_(0.0)


def construct_modulelist_from_single_module(module: torch.nn.Module, sizes: Tuple[int, ...]) -> torch.nn.Module:
    """A (nested) ModuleList of shape ``sizes`` of independent copies of ``module`` with freshly initialised Linear layers."""
    if len(sizes) == 1:
        return torch.nn.ModuleList([copy.deepcopy(module).apply(init_mlp_weights_xavier_uniform) for _ in range(sizes[0])])
    return torch.nn.ModuleList([construct_modulelist_from_single_module(module, sizes[1:]) for _ in range(sizes[0])])


def reset_module_states_post_sharding(module: torch.nn.Module) -> None:
    """Drop what modules cached from the unsharded model (e.g. the permutation of ``KTRegroupAsDict``): sharding replaces sub-modules,
    modules next to them may hold stale tensors."""
    from ..types import CacheMixin

    for sub in module.modules():
        if isinstance(sub, CacheMixin):
            sub.clear_cache()


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
