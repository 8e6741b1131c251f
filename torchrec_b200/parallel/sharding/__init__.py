"""Sharding types as route descriptions.

The reference keeps one module family per sharding type under ``torchrec/distributed/sharding/`` (``tw_sharding.py``, ``rw_sharding.py``,
``cw_sharding.py``, ``twrw_sharding.py``, ``twcw_sharding.py``, ``grid_sharding.py``, ``dp_sharding.py`` and their ``*_sequence_sharding`` /
``*_pool_sharding`` siblings), each with its own input dist, lookup and output dist classes. Here a sharding type is only a way of cutting a
table into rectangles: ``engine.shards_of`` turns a ``ParameterSharding`` of ANY type into ``TableShard``s, ``ShardedLookupEngine`` expands them to
lookup units and runs the same three kernels for all of them (``parallel/engine.py``). This package gives the per-type names
a home: ``describe(sharding_type)`` says how the type routes ids and combines results, ``units_for(...)`` returns the rectangles."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

from ..types import ParameterSharding, ShardingType


@dataclass(frozen=True)
class ShardingRoute:
    sharding_type: str
    input_route: str      # how a feature's ids reach the owner(s)
    combine: str          # how the owners' partial results form the sample's embedding
    reference_module: str


_ROUTES: Dict[str, ShardingRoute] = {
    ShardingType.DATA_PARALLEL.value: ShardingRoute("data_parallel", "none (table replicated, local lookup)", "none; dense gradient all-reduced", "sharding/dp_sharding.py"),
    ShardingType.TABLE_WISE.value: ShardingRoute("table_wise", "whole feature to the owner rank (key permutation)", "owner writes the pooled row into the sample's rank", "sharding/tw_sharding.py"),
    ShardingType.COLUMN_WISE.value: ShardingRoute("column_wise", "feature replicated to every column-shard owner", "column placement: each owner fills its column block", "sharding/cw_sharding.py"),
    ShardingType.TABLE_COLUMN_WISE.value: ShardingRoute("table_column_wise", "as column_wise, owners restricted to one host", "column placement", "sharding/twcw_sharding.py"),
    ShardingType.ROW_WISE.value: ShardingRoute("row_wise", "ids bucketed by row range, each bucket to its owner", "sum of the owners' partial pools (staged slabs + reduce)", "sharding/rw_sharding.py"),
    ShardingType.TABLE_ROW_WISE.value: ShardingRoute("table_row_wise", "row bucketing inside the table's host", "sum inside the host, then placement", "sharding/twrw_sharding.py"),
    ShardingType.GRID_SHARD.value: ShardingRoute("grid_shard", "column blocks, each row-bucketed inside a host", "sum per column block, column placement", "sharding/grid_sharding.py"),
}


def describe(sharding_type: str) -> ShardingRoute:
    return _ROUTES[ShardingType(sharding_type).value]


def all_routes() -> List[ShardingRoute]:
    return list(_ROUTES.values())


def units_for(table_idx: int, config, parameter_sharding: ParameterSharding):
    """The rectangles (``TableShard``) a plan entry cuts the table into - the only thing the engine needs to know about the type."""
    from ..engine import shards_of

    return shards_of(table_idx, config, parameter_sharding)
