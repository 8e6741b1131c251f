"""Sharding-plan report (reference planner/stats.py:148-1256): per-rank HBM/DDR, perf breakdown, imbalance."""
from __future__ import annotations

import logging
import math
import statistics
from collections import defaultdict
from typing import Any, Dict, List, Optional, Tuple, cast

from torch import nn

from ..types import ModuleSharder, ParameterSharding, ShardingPlan, ShardingType
from .types import ParameterConstraints, Perf, ShardingOption, Stats, Storage, StorageReservation, Topology
from .utils import bytes_to_gb, bytes_to_mb

logger = logging.getLogger(__name__)

MIN_WIDTH = 90


def _normalize(p: List[float]) -> List[float]:
    p_total = sum(p)
    if p_total == 0:
        return [1.0 / len(p)] * len(p)
    return [p_i / p_total for p_i in p]


def _total_variation(p: List[float]) -> float:
    k = len(p)
    if not k:
        return -1.0
    return max(abs(pi - 1.0 / k) for pi in p)


def _total_distance(p: List[float]) -> float:
    k = len(p)
    if not k:
        return -1.0
    return sum(abs(pi - 1.0 / k) for pi in p)


def _chi_divergence(p: List[float], alpha: float = 1.0) -> float:
    assert alpha >= 1
    k = len(p)
    if not k:
        return -1.0
    return sum(abs(pi - 1.0 / k) ** alpha * k ** (alpha - 1.0) for pi in p)


def _kl_divergence(p: List[float]) -> float:
    k = len(p)
    if not k:
        return -1.0
    return sum(pi * math.log(k * pi) for pi in p if pi > 0)


def _calc_max_chi_divergence(N: int, alpha: float) -> float:
    return (N - 1) ** alpha * (1.0 / N) + (N - 1) * (1.0 / N)


def _calc_max_kl_divergence(N: int) -> float:
    return math.log(N)


class EmbeddingStats(Stats):
    """Logs a table with, per rank: HBM / DDR used, estimated perf (ms), number of shards by sharding type, then
    imbalance statistics and the per-parameter placement. The formatted lines are kept in ``self._stats_table``."""

    def __init__(self) -> None:
        self._width: int = MIN_WIDTH
        self._stats_table: List[str] = []

    def log(self, sharding_plan: ShardingPlan, topology: Topology, batch_size: int, storage_reservation: StorageReservation, num_proposals: int,
            num_plans: int, run_time: float, best_plan: List[ShardingOption], constraints: Optional[Dict[str, ParameterConstraints]] = None,
            sharders: Optional[List[ModuleSharder[nn.Module]]] = None, debug: bool = True) -> None:
        shard_by_fqn = {module_name + "." + param_name: value
                        for module_name, param_dict in sharding_plan.plan.items() for param_name, value in cast(Dict[str, ParameterSharding], param_dict).items()}
        W = topology.world_size
        used_hbm = [0] * W
        used_ddr = [0] * W
        perf = [Perf() for _ in range(W)]
        stype_count: List[Dict[str, int]] = [defaultdict(int) for _ in range(W)]
        for so in best_plan:
            for shard in so.shards:
                r = cast(int, shard.rank)
                used_hbm[r] += cast(Storage, shard.storage).hbm
                used_ddr[r] += cast(Storage, shard.storage).ddr
                perf[r] += cast(Perf, shard.perf)
                stype_count[r][_abbr(so.sharding_type)] += 1
        reserved = storage_reservation.last_reserved_topology
        table: List[str] = []
        table.append("--- Planner Statistics ---")
        table.append(f"--- Evaluated {num_proposals} proposal(s), found {num_plans} possible plan(s), ran for {run_time:.2f}s ---")
        table.append(f"--- Batch size {batch_size}, world size {W}, local world size {topology.local_world_size}, device {topology.compute_device} ---")
        hdr = f"{'Rank':>5} {'HBM (GB)':>12} {'HBM %':>8} {'DDR (GB)':>10} {'Perf (ms)':>12} {'fwd c/comm':>16} {'bwd c/comm':>16}  Shards"
        table.append(hdr)
        for r in range(W):
            cap = topology.devices[r].storage.hbm or 1
            p = perf[r]
            shards = " ".join(f"{k}:{v}" for k, v in sorted(stype_count[r].items()))
            table.append(f"{r:>5} {bytes_to_gb(used_hbm[r]):>12.3f} {100.0 * used_hbm[r] / cap:>7.1f}% {bytes_to_gb(used_ddr[r]):>10.3f} {p.total:>12.4f} "
                         f"{p.fwd_compute:>7.3f}/{p.fwd_comms:<8.3f} {p.bwd_compute:>7.3f}/{p.bwd_comms:<8.3f}  {shards}")
        totals = [p.total for p in perf]
        norm = _normalize(totals)
        table.append(f"Perf: max {max(totals):.4f} ms, mean {statistics.mean(totals):.4f} ms, imbalance: total variation {_total_variation(norm):.3f}, "
                     f"KL {_kl_divergence(norm):.3f} (max {_calc_max_kl_divergence(W) if W > 1 else 0:.3f}), chi-2 {_chi_divergence(norm, 2.0):.3f}")
        hnorm = _normalize([float(h) for h in used_hbm])
        table.append(f"HBM: max {bytes_to_gb(max(used_hbm)):.3f} GB, imbalance: total variation {_total_variation(hnorm):.3f}, KL {_kl_divergence(hnorm):.3f}")
        if reserved is not None:
            table.append(f"Reserved for dense/KJT: {bytes_to_gb(topology.devices[0].storage.hbm - reserved.devices[0].storage.hbm):.3f} GB HBM per rank")
        crit = max(range(W), key=lambda r: totals[r])
        table.append(f"Critical path rank: {crit}")
        # the step is bounded by the slowest rank of each PHASE (phases are separated by device barriers), not of the sum
        phase_max = {k: max(getattr(p, k) for p in perf) for k in ("fwd_compute", "fwd_comms", "bwd_compute", "bwd_comms", "prefetch_compute")}
        table.append("Per-phase critical path (max over ranks, ms): " + ", ".join(f"{k} {v:.4f}" for k, v in phase_max.items())
                     + f"  -> sum {sum(phase_max.values()):.4f}")
        # wire volume per step and rank, from the shard geometry: ids in (input dist), pooled rows out (output dist) and the same back as grads
        ids_in, emb_out = [0.0] * W, [0.0] * W
        for so in best_plan:
            pooling = float(sum(so.input_lengths)) if so.input_lengths else 1.0
            for shard in so.shards:
                r = cast(int, shard.rank)
                frac = shard.size[0] / max(so.tensor.shape[0], 1) if so.sharding_type in ("row_wise", "table_row_wise", "grid_shard") else 1.0
                ids_in[r] += so.batch_size * W * pooling * frac * 8
                per_sample = shard.size[1] * (pooling if not so.is_pooled else 1.0)
                emb_out[r] += so.batch_size * W * per_sample * 2
        table.append(f"{'Rank':>5} {'ids in (MB)':>13} {'emb out (MB)':>14} {'grads in (MB)':>15}   (per step; bf16 rows on the wire, 8-byte ids)")
        for r in range(W):
            table.append(f"{r:>5} {bytes_to_mb(ids_in[r]):>13.2f} {bytes_to_mb(emb_out[r]):>14.2f} {bytes_to_mb(emb_out[r]):>15.2f}")
        if topology.local_world_size and W > topology.local_world_size:
            L = topology.local_world_size
            for h in range(W // L):
                rs = range(h * L, (h + 1) * L)
                table.append(f"Host {h}: HBM {bytes_to_gb(sum(used_hbm[r] for r in rs)):.2f} GB, DDR {bytes_to_gb(sum(used_ddr[r] for r in rs)):.2f} GB, "
                             f"slowest rank {max(rs, key=lambda r: totals[r])} at {max(totals[r] for r in rs):.4f} ms")
        kernels: Dict[str, int] = defaultdict(int)
        for so in best_plan:
            kernels[so.compute_kernel] += 1
        table.append("Compute kernels: " + ", ".join(f"{k} x{v}" for k, v in sorted(kernels.items())))
        top_hbm = sorted(best_plan, key=lambda so: -sum(cast(Storage, sh.storage).hbm for sh in so.shards))[:5]
        table.append("Largest tables (HBM): " + ", ".join(f"{so.name} {bytes_to_gb(sum(cast(Storage, sh.storage).hbm for sh in so.shards)):.2f} GB" for so in top_hbm))
        top_perf = sorted(best_plan, key=lambda so: -max(cast(Perf, sh.perf).total for sh in so.shards))[:5]
        table.append("Slowest shards: " + ", ".join(f"{so.name} {max(cast(Perf, sh.perf).total for sh in so.shards):.4f} ms ({_abbr(so.sharding_type)})" for so in top_perf))
        if constraints:
            table.append(f"Constraints on {len(constraints)} parameter(s): " + "; ".join(
                f"{n}: {','.join(_abbr(t) for t in (c.sharding_types or [])) or 'any'}" + (f" / {','.join(c.compute_kernels)}" if c.compute_kernels else "")
                for n, c in list(constraints.items())[:8]) + (" ..." if len(constraints) > 8 else ""))
        if debug:
            table.append(f"{'FQN':<48} {'Sharding':>9} {'Kernel':>18} {'Shards':>7} {'Rows x Dim':>18} {'Pooling':>8} {'CLF':>5} {'Perf max (ms)':>14}  Ranks")
            for so in best_plan:
                ranks = [cast(int, s.rank) for s in so.shards]
                shape = f"{so.tensor.shape[0]} x {so.tensor.shape[1]}"
                pooling = sum(so.input_lengths) if so.input_lengths else 0.0
                clf = so.cache_load_factor
                table.append(f"{so.fqn[-48:]:<48} {_abbr(so.sharding_type):>9} {so.compute_kernel:>18} {so.num_shards:>7} {shape:>18} {pooling:>8.1f} "
                             f"{(f'{clf:.2f}' if clf is not None else '-'):>5} {max(cast(Perf, s.perf).total for s in so.shards):>14.4f}  {_collapse(ranks)}")
        self._stats_table = table
        width = max(len(l) for l in table) + 4
        self._width = max(width, MIN_WIDTH)
        logger.info("#" * self._width)
        for line in table:
            logger.info("# " + line.ljust(self._width - 4) + " #")
        logger.info("#" * self._width)


class NoopEmbeddingStats(Stats):
    def log(self, *args: Any, **kwargs: Any) -> None:
        pass


def _abbr(sharding_type: str) -> str:
    return {"data_parallel": "DP", "table_wise": "TW", "column_wise": "CW", "row_wise": "RW", "table_row_wise": "TWRW",
            "table_column_wise": "TWCW", "grid_shard": "GRID"}.get(sharding_type, sharding_type)


def _collapse(ranks: List[int]) -> str:
    if not ranks:
        return ""
    if len(ranks) > 2 and ranks == list(range(ranks[0], ranks[-1] + 1)):
        return f"{ranks[0]}-{ranks[-1]}"
    return ",".join(str(r) for r in ranks)


def round_to_one_sigfig(x: float) -> str:
    """``0.0342 -> '0.03'``, ``1234 -> '1000'``: one significant figure, no exponent for ordinary magnitudes."""
    return f'{float(f"{x:.1g}"):g}'
