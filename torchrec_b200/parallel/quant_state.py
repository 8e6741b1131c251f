"""State view of sharded quantized (inference) embedding modules (reference torchrec/distributed/quant_state.py:49-637).

A sharded quantized collection holds one inference TBE per local device; serving infrastructure wants (a) a table-keyed state dict
(``embedding_bags.<table>.weight`` -> the quantized uint8 shards with their offsets) and (b) a flat description of where every shard of
every table lives (``sharded_tbes_weights_spec``) to stream weight updates into a running server."""
from __future__ import annotations

import re
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from .shards_wrapper import LocalShardsWrapper

_SHARD_NAME = re.compile(r"^(?P<table>.+)_(?P<row>\d+)_(?P<col>\d+)$")


@dataclass
class WeightSpec:
    fqn: str                 # "<module>.embedding_bags.<table>.weight" of the UNSHARDED model
    shard_offsets: List[int]  # [row, col]
    shard_sizes: List[int]    # [rows, cols] (logical elements, not bytes)
    sharding_type: Optional[str]
    device: str = ""
    data_type: str = ""


def _shards_of(module: nn.Module) -> List[Tuple[str, int, int, int, int, Any, torch.Tensor, str]]:
    """(table, row_off, col_off, rows, cols, data_type, quantized rows [rows, row_bytes], device) for every local shard."""
    out = []
    for tbe in getattr(module, "_tbes", []):
        if not hasattr(tbe, "split_embedding_weights"):
            continue
        for (name, rows, cols, dt), w in zip(tbe.embedding_specs, tbe.split_embedding_weights()):
            m = _SHARD_NAME.match(str(name))
            table, r0, c0 = (m.group("table"), int(m.group("row")), int(m.group("col"))) if m else (str(name), 0, 0)
            out.append((table, r0, c0, int(rows), int(cols), dt, w, str(w.device)))
    return out


def _sharding_type(shards: List[Tuple[int, int]]) -> str:
    if len(shards) <= 1:
        return "table_wise"
    rows_vary = len({r for r, _ in shards}) > 1
    cols_vary = len({c for _, c in shards}) > 1
    if rows_vary and cols_vary:
        return "grid_shard"
    return "row_wise" if rows_vary else "column_wise"


def _weight_prefix(module: nn.Module) -> str:
    return "embeddings" if type(module).__name__.endswith("EmbeddingCollection") and "Bag" not in type(module).__name__ else "embedding_bags"


class ShardedQuantEmbeddingModuleState:
    """Mixin / helper: ``sharded_state_dict(module)`` -> table-keyed quantized state with shard offsets."""

    @staticmethod
    def sharded_state_dict(module: nn.Module, prefix: str = "") -> "OrderedDict[str, Any]":
        by_table: Dict[str, List[Tuple[int, int, torch.Tensor]]] = {}
        for table, r0, c0, _rows, _cols, _dt, w, _dev in _shards_of(module):
            by_table.setdefault(table, []).append((r0, c0, w))
        kind = _weight_prefix(module)
        out: "OrderedDict[str, Any]" = OrderedDict()
        for table, shards in by_table.items():
            key = f"{prefix}{kind}.{table}.weight"
            if len(shards) == 1 and shards[0][0] == 0 and shards[0][1] == 0:
                out[key] = shards[0][2]
            else:
                shards.sort(key=lambda s: (s[0], s[1]))
                out[key] = LocalShardsWrapper([w for _, _, w in shards], [(r0, c0) for r0, c0, _ in shards])
        return out


def sharded_tbes_weights_spec(sharded_model: nn.Module, virtual_table_name_to_bucket_lengths: Optional[Dict[str, List[int]]] = None) -> Dict[str, WeightSpec]:
    """``"<fqn>.tbes.<device idx>.<shard idx>.weight" -> WeightSpec`` for every shard of every sharded quantized collection in the model."""
    ret: Dict[str, WeightSpec] = {}
    for fqn, module in sharded_model.named_modules():
        if not hasattr(module, "_tbes") or not any(hasattr(t, "split_embedding_weights") for t in module._tbes):
            continue
        shards = _shards_of(module)
        per_table: Dict[str, List[Tuple[int, int]]] = {}
        for table, r0, c0, *_ in shards:
            per_table.setdefault(table, []).append((r0, c0))
        kind = _weight_prefix(module)
        i = 0
        for di, tbe in enumerate(module._tbes):
            if not hasattr(tbe, "split_embedding_weights"):
                continue
            for k in range(len(tbe.embedding_specs)):
                table, r0, c0, rows, cols, dt, w, dev = shards[i]
                i += 1
                key = f"{fqn + '.' if fqn else ''}tbes.{di}.{k}.weight"
                ret[key] = WeightSpec(fqn=f"{fqn + '.' if fqn else ''}{kind}.{table}.weight", shard_offsets=[r0, c0], shard_sizes=[rows, cols],
                                      sharding_type=_sharding_type(per_table[table]), device=dev, data_type=str(getattr(dt, "name", dt)))
    return ret


def get_param_id_from_type(is_sqebc: bool, is_sqmcec: bool, is_sfpebc: bool) -> str:
    """Attribute path of the per-table weights inside a sharded quantized module's state dict: bags, managed-collision sequence
    collection, feature-processed bags, or (default) plain sequence collection."""
    if is_sqebc:
        return "embedding_bags"
    if is_sqmcec:
        return "_embedding_module.embeddings"
    if is_sfpebc:
        return "_embedding_bag_collection.embedding_bags"
    return "embeddings"


def get_bucket_offsets_per_virtual_table(grouped_embedding_config: List[Any], virtual_table_name_to_bucket_lengths: Dict[str, List[int]]) -> Dict[str, List[int]]:
    """For every virtual (key-value) table: the index of the first bucket of each of its row shards, from the shard metadata of the
    grouped configs (all groups must agree on a table's shards) and the number of buckets the table was trained with."""
    from collections import defaultdict

    from .utils import get_bucket_metadata_from_shard_metadata

    shards_of: Dict[str, List[Any]] = defaultdict(list)
    for config in grouped_embedding_config:
        for table in config.embedding_tables:
            meta = getattr(table, "global_metadata", None)
            assert meta is not None and meta.shards_metadata is not None, f"Table: {table.name} doesn't have global metadata in grouped embedding config"
            if table.name in virtual_table_name_to_bucket_lengths:
                if table.name in shards_of:
                    assert shards_of[table.name] == meta.shards_metadata, f"Virtual table: {table.name} should have same global metadata across all grouped embedding configs"
                else:
                    shards_of[table.name] = meta.shards_metadata
    return {name: get_bucket_metadata_from_shard_metadata(shards, len(virtual_table_name_to_bucket_lengths[name])).bucket_offsets_per_shard for name, shards in shards_of.items()}


def post_state_dict_hook(module: nn.Module, destination: Dict[str, torch.Tensor], prefix: str, _local_metadata: Dict[str, Any], tables_weights_prefix: str) -> None:
    """State-dict hook of a sharded quantized module: per table ``<prefix><tables_weights_prefix>.<table>.weight`` (+ ``weight_qscale`` /
    ``weight_qbias`` when scale and bias are split) from the module's per-table sharded views."""
    for name, t in getattr(module, "_table_name_to_sharded_tensor", {}).items():
        destination[f"{prefix}{tables_weights_prefix}.{name}.weight"] = t
    for sfx, sharded, lists in (("weight_qscale", "_table_name_to_sharded_tensor_qscale", "_table_name_to_tensors_list_qscale"),
                                ("weight_qbias", "_table_name_to_sharded_tensor_qbias", "_table_name_to_tensors_list_qbias")):
        for name, t in getattr(module, sharded, {}).items():
            destination[f"{prefix}{tables_weights_prefix}.{name}.{sfx}"] = t
        for name, ts in getattr(module, lists, {}).items():
            for i, t in enumerate(ts):
                destination[f"{prefix}{tables_weights_prefix}.{name}.{sfx}.{i}"] = t
