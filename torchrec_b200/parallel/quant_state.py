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


def sharded_tbes_weights_spec(sharded_model: nn.Module) -> Dict[str, WeightSpec]:
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
