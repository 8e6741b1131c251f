"""Sharding-plan construction helpers (reference torchrec/distributed/sharding_plan.py:48-966).

Shard size/offset calculators for every ShardingType, ``ParameterSharding`` generators
(``table_wise(rank=..)``, ``row_wise()``, ``column_wise(ranks=..)``, ``table_row_wise(host_index=..)``,
``grid_shard(host_indexes=..)``, ``data_parallel()``) and ``construct_module_sharding_plan``.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple, Type, Union

import torch
from torch import nn

from .embedding_types import EmbeddingComputeKernel
from .types import (
    EmbeddingModuleShardingPlan,
    EnumerableShardingSpec,
    ModuleSharder,
    ParameterSharding,
    ShardingType,
    ShardMetadata,
)

MIN_CW_DIM: int = 128


def get_default_sharders() -> List[ModuleSharder[nn.Module]]:
    from .embedding import EmbeddingCollectionSharder
    from .embeddingbag import EmbeddingBagCollectionSharder
    from .extra_sharders import default_extra_sharders

    return [EmbeddingBagCollectionSharder(), EmbeddingCollectionSharder(), *default_extra_sharders()]


def get_module_to_default_sharders() -> Dict[Type[nn.Module], ModuleSharder[nn.Module]]:
    return {s.module_type: s for s in get_default_sharders()}


def placement(compute_device: str, rank: int, local_size: int) -> str:
    """Placement string ``rank:<r>/<device>[:<local_rank>]``."""
    s = f"rank:{rank}/{compute_device}"
    if compute_device in ("cuda", "mtia"):
        s += f":{rank % local_size}"
    return s


def placement_rank(pl) -> int:
    """Rank of a placement (string or torch.distributed._remote_device)."""
    if hasattr(pl, "rank"):
        r = pl.rank()
        if r is not None:
            return r
    s = str(pl)
    return int(s.split("/")[0].split(":")[1])


# ---- size / offset calculators ------------------------------------------------------------------
def _rw_sizes_offsets(rows: int, num_devices: int, cols: int) -> Tuple[List[List[int]], List[List[int]]]:
    block = (rows + num_devices - 1) // num_devices
    last_rank = rows // block if block > 0 else 0
    last_block = rows - block * last_rank
    sizes, offs, o = [], [], 0
    for r in range(num_devices):
        n = block if r < last_rank else (last_block if r == last_rank else 0)
        sizes.append([n, cols])
        offs.append([o, 0])
        o += n
    return sizes, offs


def _uneven_rw_sizes_offsets(rows: int, cols: int, size_per_rank: List[int]) -> Tuple[List[List[int]], List[List[int]]]:
    assert sum(size_per_rank) == rows, f"uneven row-wise sizes {size_per_rank} must add up to {rows}"
    sizes, offs, o = [], [], 0
    for n in size_per_rank:
        sizes.append([n, cols])
        offs.append([o, 0])
        o += n
    return sizes, offs


def _cw_block_size(cols: int, col_wise_shard_dim: Optional[int]) -> int:
    if col_wise_shard_dim:
        return col_wise_shard_dim
    if cols >= MIN_CW_DIM * 2:
        # split into MIN_CW_DIM wide shards (rounded to a multiple of 4 for the vectorised kernels)
        return MIN_CW_DIM
    return cols


def _cw_sizes_offsets(cols: int, rows: int, col_wise_shard_dim: Optional[int] = None) -> Tuple[List[List[int]], List[List[int]]]:
    block = min(_cw_block_size(cols, col_wise_shard_dim), cols)
    if block % 4 != 0:
        raise ValueError(f"column-wise shard dim must be a multiple of 4, got {block}")
    n, residual = divmod(cols, block)
    sizes = [[rows, block] for _ in range(n)]
    if residual:
        sizes[-1][1] += residual
    offs, o = [], 0
    for s in sizes:
        offs.append([0, o])
        o += s[1]
    return sizes, offs


def _twrw_sizes_offsets(rows: int, cols: int, local_world_size: int) -> Tuple[List[List[int]], List[List[int]]]:
    return _rw_sizes_offsets(rows, local_world_size, cols)


def _grid_sizes_offsets(rows: int, cols: int, local_world_size: int, col_wise_shard_dim: Optional[int]) -> Tuple[List[List[int]], List[List[int]], int]:
    cw_sizes, cw_offs = _cw_sizes_offsets(cols, rows, col_wise_shard_dim)
    sizes, offs = [], []
    for cs, co in zip(cw_sizes, cw_offs):
        rs, ro = _rw_sizes_offsets(rows, local_world_size, cs[1])
        for s, o in zip(rs, ro):
            sizes.append(s)
            offs.append([o[0], co[1]])
    return sizes, offs, len(cw_sizes)


def calculate_shard_sizes_and_offsets(
    tensor: torch.Tensor,
    world_size: int,
    local_world_size: int,
    sharding_type: str,
    col_wise_shard_dim: Optional[int] = None,
    device_memory_sizes: Optional[List[int]] = None,
    num_buckets: Optional[int] = None,
) -> Tuple[List[List[int]], List[List[int]]]:
    """Shard sizes / offsets of a [rows, cols] table under a sharding type. ``num_buckets`` (row-wise only; managed-collision tables):
    the rows are cut into that many equal buckets and whole buckets are dealt to the ranks, the first ``num_buckets % world_size`` ranks
    taking one more - a bucket never straddles two shards."""
    rows, cols = tensor.shape
    if sharding_type == ShardingType.DATA_PARALLEL.value:
        return [[rows, cols]] * world_size, [[0, 0]] * world_size
    if sharding_type == ShardingType.TABLE_WISE.value:
        return [[rows, cols]], [[0, 0]]
    if sharding_type == ShardingType.ROW_WISE.value:
        if num_buckets:
            assert rows % num_buckets == 0, "hash_size must be divisible by num_buckets"
            bucket = rows // num_buckets
            per, extra = num_buckets // world_size, num_buckets % world_size
            sizes = [[bucket * (per + (1 if r < extra else 0)), cols] for r in range(world_size)]
            offs, o = [], 0
            for s in sizes:
                offs.append([o, 0])
                o += s[0]
            return sizes, offs
        return _rw_sizes_offsets(rows, world_size, cols)
    if sharding_type == ShardingType.TABLE_ROW_WISE.value:
        return _twrw_sizes_offsets(rows, cols, local_world_size)
    if sharding_type in (ShardingType.COLUMN_WISE.value, ShardingType.TABLE_COLUMN_WISE.value):
        return _cw_sizes_offsets(cols, rows, col_wise_shard_dim)
    if sharding_type == ShardingType.GRID_SHARD.value:
        s, o, _ = _grid_sizes_offsets(rows, cols, local_world_size, col_wise_shard_dim)
        return s, o
    raise ValueError(f"Unrecognized or unsupported sharding type provided: {sharding_type}")


def _spec(sizes: List[List[int]], offs: List[List[int]], ranks: List[int], device_type: str, local_size: int) -> EnumerableShardingSpec:
    return EnumerableShardingSpec([
        ShardMetadata(shard_sizes=list(s), shard_offsets=list(o), placement=placement(device_type, r, local_size))
        for s, o, r in zip(sizes, offs, ranks)
    ])


def _get_parameter_sharding(
    param: Union[nn.Parameter, torch.Tensor],
    sharding_type: str,
    size_offset_ranks: List[Tuple[List[int], List[int], int]],
    local_size: int,
    device_type: str,
    sharder: ModuleSharder[nn.Module],
    placements: Optional[List[str]] = None,
    compute_kernel: Optional[str] = None,
) -> ParameterSharding:
    return ParameterSharding(
        sharding_spec=None if sharding_type == ShardingType.DATA_PARALLEL.value else EnumerableShardingSpec([
            ShardMetadata(shard_sizes=list(size), shard_offsets=list(offset),
                          placement=placements[i] if placements else placement(device_type, rank, local_size))
            for i, (size, offset, rank) in enumerate(size_offset_ranks)
        ]),
        sharding_type=sharding_type,
        compute_kernel=compute_kernel if compute_kernel else _get_compute_kernel(sharder, param, sharding_type, device_type),
        ranks=[rank for (_, _, rank) in size_offset_ranks],
    )


def _get_compute_kernel(sharder, param, sharding_type: str, device_type: str) -> str:
    kernels = [k for k in sharder.compute_kernels(sharding_type, device_type)]
    if sharding_type == ShardingType.DATA_PARALLEL.value:
        pref = [EmbeddingComputeKernel.DENSE.value]
    elif hasattr(param, "_in_backward_optimizers") or hasattr(param, "_optimizer_classes"):
        pref = [EmbeddingComputeKernel.FUSED.value, EmbeddingComputeKernel.DENSE.value]
    else:
        pref = [EmbeddingComputeKernel.FUSED.value, EmbeddingComputeKernel.DENSE.value]
    for k in pref:
        if k in kernels:
            return k
    return kernels[0]


ParameterShardingGenerator = Callable[[nn.Parameter, int, int, str, ModuleSharder[nn.Module]], ParameterSharding]


def data_parallel() -> ParameterShardingGenerator:
    def gen(param, local_size, world_size, device_type, sharder) -> ParameterSharding:
        sizes, offs = calculate_shard_sizes_and_offsets(param, world_size, local_size, ShardingType.DATA_PARALLEL.value)
        return _get_parameter_sharding(param, ShardingType.DATA_PARALLEL.value, [(s, o, r) for r, (s, o) in enumerate(zip(sizes, offs))],
                                       local_size, device_type, sharder)

    return gen


def table_wise(rank: int, device: Optional[str] = None, compute_kernel: Optional[str] = None) -> ParameterShardingGenerator:
    def gen(param, local_size, world_size, device_type, sharder) -> ParameterSharding:
        sizes, offs = calculate_shard_sizes_and_offsets(param, world_size, local_size, ShardingType.TABLE_WISE.value)
        dt = device if device is not None else device_type
        return _get_parameter_sharding(param, ShardingType.TABLE_WISE.value, [(sizes[0], offs[0], rank)], local_size, dt, sharder,
                                       compute_kernel=compute_kernel)

    return gen


def row_wise(sizes_placement: Optional[Tuple[List[int], Union[str, List[str]]]] = None, compute_kernel: Optional[str] = None) -> ParameterShardingGenerator:
    """Even row-wise over all ranks, or uneven with explicit ``(sizes, device or placements)``."""

    def gen(param, local_size, world_size, device_type, sharder) -> ParameterSharding:
        placements = None
        if sizes_placement is None:
            sizes, offs = calculate_shard_sizes_and_offsets(param, world_size, local_size, ShardingType.ROW_WISE.value)
        else:
            sizes, offs = _uneven_rw_sizes_offsets(param.shape[0], param.shape[1], sizes_placement[0])
            if isinstance(sizes_placement[1], list):
                placements = [placement(d, r, world_size) for r, d in enumerate(sizes_placement[1])]
            else:
                placements = [placement(sizes_placement[1], r, local_size) for r in range(len(sizes))]
        return _get_parameter_sharding(param, ShardingType.ROW_WISE.value, [(s, o, r) for r, (s, o) in enumerate(zip(sizes, offs))],
                                       local_size, device_type, sharder, placements=placements, compute_kernel=compute_kernel)

    return gen


def column_wise(ranks: Optional[List[int]] = None, size_per_rank: Optional[List[int]] = None, compute_kernel: Optional[str] = None,
                device_types: Optional[List[str]] = None) -> ParameterShardingGenerator:
    """Split the embedding dim over ``ranks`` (evenly) or by explicit ``size_per_rank``. ``device_types``: one device type per column shard
    (heterogeneous inference: some shards in host memory) instead of the module's device type for all."""

    def gen(param, local_size, world_size, device_type, sharder) -> ParameterSharding:
        rows, cols = param.shape
        if size_per_rank is not None:
            assert sum(size_per_rank) == cols, "size_per_rank must add up to the embedding dim"
            rk = ranks if ranks is not None else list(range(len(size_per_rank)))
            sizes = [[rows, s] for s in size_per_rank]
        else:
            assert ranks is not None, "column_wise needs ranks or size_per_rank"
            rk = ranks
            if cols % len(rk) != 0:
                raise ValueError(f"column dim of {cols} cannot be evenly divided across {rk}")
            sizes = [[rows, cols // len(rk)] for _ in rk]
        offs, o = [], 0
        for s in sizes:
            offs.append([0, o])
            o += s[1]
        placements = None
        if device_types is not None:
            assert len(device_types) == len(rk), "device_types must have one entry per column shard"
            index: Dict[str, int] = {}
            placements = []
            for dt, r in zip(device_types, rk):
                placements.append(placement_helper(dt, index.get(dt, 0), r))
                index[dt] = index.get(dt, 0) + 1
        return _get_parameter_sharding(param, ShardingType.COLUMN_WISE.value, list(zip(sizes, offs, rk)), local_size, device_type, sharder,
                                       placements=placements, compute_kernel=compute_kernel)

    return gen


def table_row_wise(host_index: int, compute_kernel: Optional[str] = None) -> ParameterShardingGenerator:
    def gen(param, local_size, world_size, device_type, sharder) -> ParameterSharding:
        sizes, offs = calculate_shard_sizes_and_offsets(param, world_size, local_size, ShardingType.TABLE_ROW_WISE.value)
        ranks = [host_index * local_size + i for i in range(local_size)]
        return _get_parameter_sharding(param, ShardingType.TABLE_ROW_WISE.value, list(zip(sizes, offs, ranks)), local_size, device_type, sharder,
                                       compute_kernel=compute_kernel)

    return gen


def grid_shard(host_indexes: List[int], compute_kernel: Optional[str] = None) -> ParameterShardingGenerator:
    """Column shards placed on the hosts ``host_indexes``; each row-split over that host's ranks."""

    def gen(param, local_size, world_size, device_type, sharder) -> ParameterSharding:
        rows, cols = param.shape
        assert cols % len(host_indexes) == 0, "columns must divide evenly over the hosts"
        dim = cols // len(host_indexes)
        sizes, offs, ranks = [], [], []
        for ci, h in enumerate(host_indexes):
            rs, ro = _rw_sizes_offsets(rows, local_size, dim)
            for i, (s, o) in enumerate(zip(rs, ro)):
                sizes.append(s)
                offs.append([o[0], ci * dim])
                ranks.append(h * local_size + i)
        return _get_parameter_sharding(param, ShardingType.GRID_SHARD.value, list(zip(sizes, offs, ranks)), local_size, device_type, sharder,
                                       compute_kernel=compute_kernel)

    return gen


def apply_to_all(module: nn.Module, parameter_sharding_generator: ParameterShardingGenerator, sharder: Optional[ModuleSharder[nn.Module]] = None) -> Dict[str, ParameterShardingGenerator]:
    if sharder is None:
        sharder = get_module_to_default_sharders().get(type(module), None)
    assert sharder is not None, f"no default sharder for {type(module)}"
    return {name: parameter_sharding_generator for name in sharder.shardable_parameters(module)}


def construct_module_sharding_plan(
    module: nn.Module,
    per_param_sharding: Dict[str, ParameterShardingGenerator],
    sharder: Optional[ModuleSharder[nn.Module]] = None,
    local_size: Optional[int] = None,
    world_size: Optional[int] = None,
    device_type: Optional[str] = None,
) -> EmbeddingModuleShardingPlan:
    """Build an ``EmbeddingModuleShardingPlan`` from per-table generators.

    Example::

        plan = construct_module_sharding_plan(ebc, {"t0": table_wise(rank=0), "t1": row_wise()})
    """
    import torch.distributed as dist

    from .comm import get_local_size

    if device_type is None:
        device_type = "cuda" if torch.cuda.is_available() else "cpu"
    if sharder is None:
        sharder = get_module_to_default_sharders().get(type(module), None)
    assert sharder is not None, f"Could not find a valid sharder type for {type(module)}"
    assert isinstance(module, sharder.module_type), f"Incorrect sharder {type(sharder)} for module {type(module)}"
    shardable = sharder.shardable_parameters(module)
    assert shardable.keys() == per_param_sharding.keys(), "per_param_sharding_config doesn't match the shardable parameters of the module," \
        f" got {list(shardable.keys())} != {list(per_param_sharding.keys())}."
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if local_size is None:
        local_size = get_local_size(world_size)
    plan = EmbeddingModuleShardingPlan()
    for name, gen in per_param_sharding.items():
        plan[name] = gen(shardable[name], local_size, world_size, device_type, sharder)
    return plan


def placement_helper(device_type: str, index: int = 0, rank: int = 0) -> str:
    """Placement string of a shard: ``rank:<rank>/<device>:<index>``; cpu shards all live on rank 0's host."""
    return f"rank:0/{device_type}" if device_type == "cpu" else f"rank:{rank}/{device_type}:{index}"


def get_sharding_constructor_from_type(sharding_type: ShardingType) -> Callable[..., ParameterShardingGenerator]:
    """The ``construct_module_sharding_plan`` helper (``table_wise``, ``row_wise``, ...) that makes a placement of this type."""
    return {ShardingType.TABLE_WISE: table_wise, ShardingType.ROW_WISE: row_wise, ShardingType.COLUMN_WISE: column_wise, ShardingType.TABLE_ROW_WISE: table_row_wise,
            ShardingType.GRID_SHARD: grid_shard, ShardingType.DATA_PARALLEL: data_parallel}[sharding_type]
