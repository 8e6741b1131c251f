"""Dynamic re-sharding: move table shards of a live sharded module to a new placement
(reference torchrec/distributed/sharding/dynamic_sharding.py:1-1100, model_parallel.py:793 ``reshard``).

The reference moves whole shards with all-to-all of flattened tensors and only supports table-wise / column-wise moves.
Here the old and the new layout are both lists of rectangles (``TableShard``), so re-sharding is rectangle intersection:
every (old shard ∩ new shard) block is either a local copy or one point-to-point send/recv (NCCL over NVLink / gloo),
for the weights and for every optimizer state — any sharding type to any sharding type."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..engine import TableShard, shards_of
from ..types import ParameterSharding, ShardingEnv, ShardingType


class _Stub(nn.Module):
    """What a sharder needs from the unsharded module: configs, weighting, optimizer specs (no storage)."""

    def __init__(self, old) -> None:
        super().__init__()
        self._old = old
        self._trb_opt_specs = dict(getattr(old, "_opt_specs", {}))

    def is_weighted(self) -> bool:
        return self._old.is_weighted()

    def embedding_bag_configs(self):
        return self._old.embedding_bag_configs()

    def embedding_configs(self):
        return self._old.embedding_configs()

    def need_indices(self) -> bool:
        return self._old.need_indices() if hasattr(self._old, "need_indices") else False

    def embedding_dim(self) -> int:
        return self._old.embedding_dim()

    def embedding_names_by_table(self):
        return self._old.embedding_names_by_table()


def _intersect(a: TableShard, b: TableShard) -> Optional[Tuple[int, int, int, int]]:
    r0, r1 = max(a.row_off, b.row_off), min(a.row_off + a.rows, b.row_off + b.rows)
    c0, c1 = max(a.col_off, b.col_off), min(a.col_off + a.cols, b.col_off + b.cols)
    if r0 >= r1 or c0 >= c1:
        return None
    return r0, r1, c0, c1


def _all_shards(module) -> List[TableShard]:
    cfgs = module.embedding_bag_configs() if hasattr(module, "embedding_bag_configs") else module.embedding_configs()
    res: List[TableShard] = []
    for ti, c in enumerate(cfgs):
        ps = module._plan[c.name]
        if ps.sharding_type == ShardingType.DATA_PARALLEL.value:
            continue
        res.extend(shards_of(ti, c, ps))
    return res


@torch.no_grad()
def transfer_state(old, new, env: ShardingEnv, changed: Optional[List[str]] = None) -> int:
    """Copy weights + optimizer states of ``old`` into ``new`` (same tables, different placement). Returns bytes moved
    over the wire by this rank. Collective over ``env.process_group``."""
    rank = env.rank
    pg = env.process_group
    old_views = {(s.name, s.row_off, s.col_off): (s, w, st) for s, w, st, _ in old.engine.local_shard_views()} if old.engine is not None else {}
    new_views = {(s.name, s.row_off, s.col_off): (s, w, st) for s, w, st, _ in new.engine.local_shard_views()} if new.engine is not None else {}
    old_shards, new_shards = _all_shards(old), _all_shards(new)
    ops: List[dist.P2POp] = []
    pending: List[Tuple[torch.Tensor, torch.Tensor]] = []  # (recv buffer, destination view)
    keep: List[torch.Tensor] = []
    moved = 0
    for ns in new_shards:
        for os_ in old_shards:
            if os_.name != ns.name:
                continue
            box = _intersect(os_, ns)
            if box is None:
                continue
            r0, r1, c0, c1 = box
            src_local, dst_local = os_.rank == rank, ns.rank == rank
            if not src_local and not dst_local:
                continue
            _, ow, ost = old_views[(os_.name, os_.row_off, os_.col_off)] if src_local else (None, None, {})
            _, nw, nst = new_views[(ns.name, ns.row_off, ns.col_off)] if dst_local else (None, None, {})
            # state names are decided by the optimizer, identical on both sides
            names = sorted(ost.keys()) if src_local else sorted(nst.keys())
            blocks = [("w", None)] + [("s", n) for n in names]
            for kind, n in blocks:
                if kind == "w":
                    src = ow[r0 - os_.row_off : r1 - os_.row_off, c0 - os_.col_off : c1 - os_.col_off] if src_local else None
                    dst = nw[r0 - ns.row_off : r1 - ns.row_off, c0 - ns.col_off : c1 - ns.col_off] if dst_local else None
                else:
                    sv = ost[n] if src_local else None
                    dv = nst[n] if dst_local else None
                    rowwise = (sv if src_local else dv).dim() == 1
                    if rowwise:
                        # one value per row and column shard: the new shard inherits it from the old shard holding its first column
                        if not (os_.col_off <= ns.col_off < os_.col_off + os_.cols):
                            continue
                        src = sv[r0 - os_.row_off : r1 - os_.row_off] if src_local else None
                        dst = dv[r0 - ns.row_off : r1 - ns.row_off] if dst_local else None
                    else:
                        src = sv[r0 - os_.row_off : r1 - os_.row_off, c0 - os_.col_off : c1 - os_.col_off] if src_local else None
                        dst = dv[r0 - ns.row_off : r1 - ns.row_off, c0 - ns.col_off : c1 - ns.col_off] if dst_local else None
                if src_local and dst_local:
                    dst.copy_(src)
                elif src_local:
                    buf = src.contiguous()
                    keep.append(buf)
                    ops.append(dist.P2POp(dist.isend, buf, dist.get_global_rank(pg, ns.rank) if pg is not None else ns.rank, group=pg))
                    moved += buf.numel() * buf.element_size()
                else:
                    buf = torch.empty(dst.shape, dtype=dst.dtype, device=dst.device)
                    ops.append(dist.P2POp(dist.irecv, buf, dist.get_global_rank(pg, os_.rank) if pg is not None else os_.rank, group=pg))
                    pending.append((buf, dst))
                    moved += buf.numel() * buf.element_size()
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, dst in pending:
        dst.copy_(buf)
    # data-parallel tables: replicated, copy locally
    if getattr(old, "_dp_tables", None) and getattr(new, "_dp_tables", None):
        oldw = dict(zip([old._embedding_bag_configs[ti].name for ti in old._dp_tables], old._dp_tbe.split_embedding_weights()))
        for ti, w in zip(new._dp_tables, new._dp_tbe.split_embedding_weights()):
            name = new._embedding_bag_configs[ti].name
            if name in oldw:
                w.copy_(oldw[name])
    return moved


def reshard_module(old, changed_shard_to_params: Dict[str, ParameterSharding], env: ShardingEnv, device: torch.device, sharder_map: Dict[Any, Any]):
    """Build the replacement of the sharded module ``old`` under the updated plan and move all state into it."""
    new_plan: Dict[str, ParameterSharding] = dict(old._plan)
    new_plan.update(changed_shard_to_params)
    kwargs = dict(env=env, fused_params=getattr(old, "_fused_params", None), device=device, qcomm_codecs_registry=old.qcomm_codecs_registry)
    stub = _Stub(old)
    new = type(old)(stub, new_plan, **kwargs)
    transfer_state(old, new, env, list(changed_shard_to_params.keys()))
    # carry the scalar optimizer state (step counters, learning rate) over
    # (groups without local shards hold an nn.Identity placeholder; a rank may own no table before and some after the move, so the
    # scalars are taken from whichever rank has them)
    real = lambda eng: [t for t in eng._tbes if hasattr(t, "hyper_host")] if eng is not None else []  # noqa: E731
    mine = real(old.engine)
    hyper = list(mine[0].hyper_host) if mine else None  # python list of floats (lr, eps, betas, weight decay, step, ...)
    pg = env.process_group
    if pg is not None and dist.is_initialized() and dist.get_world_size(pg) > 1:
        gathered: List[Any] = [None] * dist.get_world_size(pg)
        dist.all_gather_object(gathered, hyper, group=pg)
        if hyper is None:
            hyper = next((h for h in gathered if h is not None), None)
    if hyper is not None:
        for t in real(new.engine):
            t.hyper_host[:] = hyper
            t._push_hyper()
    new.train(old.training)
    return new


def output_sharding_plan_delta(old_plan: Dict[str, ParameterSharding], new_plan: Dict[str, ParameterSharding], return_data_volume: bool = False):
    """Tables whose placement differs between two module plans (+ optionally the bytes that would move)."""
    delta = {k: v for k, v in new_plan.items() if k not in old_plan or old_plan[k].ranks != v.ranks or old_plan[k].sharding_type != v.sharding_type
             or [s.shard_offsets for s in (old_plan[k].sharding_spec.shards if old_plan[k].sharding_spec else [])] != [s.shard_offsets for s in (v.sharding_spec.shards if v.sharding_spec else [])]}
    if not return_data_volume:
        return delta
    vol = 0
    for v in delta.values():
        if v.sharding_spec is not None:
            vol += sum(s.shard_sizes[0] * s.shard_sizes[1] * 4 for s in v.sharding_spec.shards)
    return delta, vol


# ---- plan deltas in the reference's shapes ----------------------------------------------------------------------------------------------------------------
def output_sharding_plan_delta_single(old_plan: Dict[str, ParameterSharding], new_plan: Dict[str, ParameterSharding], return_data_volume: bool = False):
    """``(megabytes that would move, {table: new ParameterSharding})`` for the tables whose placement differs between two plans of one
    module (same tables in both). What ``DistributedModelParallel.reshard`` takes as ``changed_shard_to_params``."""
    import copy

    from ..types import EmbeddingModuleShardingPlan

    assert len(old_plan) == len(new_plan), "both plans must cover the same tables"
    delta = output_sharding_plan_delta(old_plan, new_plan)
    diff = EmbeddingModuleShardingPlan({k: copy.deepcopy(v) for k, v in delta.items()})
    volume = 0.0
    if return_data_volume:
        for v in diff.values():
            if v.sharding_spec is not None:
                volume += sum(s.shard_sizes[0] * s.shard_sizes[1] * 4 / (1024 * 1024) for s in v.sharding_spec.shards)  # float rows
    return volume, diff


def output_sharding_plans_delta(old_plan: Dict[str, Dict[str, ParameterSharding]], new_plan: Dict[str, Dict[str, ParameterSharding]], return_data_volume: bool = False):
    """Per module fqn (the layout of ``ShardingPlan.plan``): ``output_sharding_plan_delta_single`` of its old and new plan."""
    out = {}
    for key, plan in old_plan.items():
        assert key in new_plan, f"module {key} is missing from the new plan"
        out[key] = output_sharding_plan_delta_single(plan, new_plan[key], return_data_volume)
    return out


def update_module_sharding_plan(module, changed_sharding_params: Dict[str, ParameterSharding], sharding_type_to_sharding_infos: Optional[Dict[str, List[Any]]] = None) -> None:
    """Record new placements in a sharded module's plan (``module_sharding_plan`` / ``_plan``) and in the matching sharding infos."""
    plan = getattr(module, "module_sharding_plan", None)
    if plan is None:
        plan = getattr(module, "_plan", None)
    if plan is None:
        raise RuntimeError("Module does not have a module_sharding_plan attribute")
    for name, param in changed_sharding_params.items():
        plan[name] = param
        for info in (sharding_type_to_sharding_infos or {}).get(param.sharding_type, []):
            if info.embedding_config.name == name:
                info.param_sharding = param


def move_sharded_tensors_to_cpu(state_dict: Dict[str, Any]) -> Dict[str, Any]:
    """Move the local shards of every ShardedTensor in a (nested) state dict to host memory - frees HBM while a re-shard or a checkpoint
    holds a second copy of the state."""
    def walk(item: Any) -> Any:
        if hasattr(item, "local_shards") and callable(item.local_shards):
            for shard in item.local_shards():
                if shard.tensor.device.type == "cuda":
                    shard.tensor = shard.tensor.cpu()
            return item
        if isinstance(item, dict):
            return {k: walk(v) for k, v in item.items()}
        if isinstance(item, list):
            return [walk(v) for v in item]
        if isinstance(item, tuple):
            return tuple(walk(v) for v in item)
        return item

    out = walk(state_dict)
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return out
