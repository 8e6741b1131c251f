"""Route-driven sharded lookup engine.

Instead of one Python module zoo per sharding type (reference ``sharding/*.py``), every
model-parallel sharding type is expressed with one notion, the *lookup unit*:

    unit = (feature f, table shard = rows [row_lo,row_hi) x cols [col_lo,col_hi) on rank r)

    TABLE_WISE        1 unit per feature                      (full rows, full cols)
    COLUMN_WISE/TWCW  n_c units per feature                   (full rows, column slices)
    ROW_WISE/TWRW     n_r units per feature                   (row ranges, full cols)
    GRID_SHARD        n_c x n_r units per feature

* input dist  : every id of feature f is routed to the unit(s) whose row range holds it (replicated
                over column slices); one KJT all-to-all moves all units of all sharding types.
* lookup      : each rank runs its table-batched kernel over its units for the global batch.
* output dist : one pooled all-to-all returns ``[B_local, sum(unit dims)]``; units of the same
                (feature, column slice) are *summed* at the destination (this is the row-wise
                reduce-scatter, done as all-to-all + local add — the NVSwitch-friendly schedule), and
                column slices are placed at their offset in the feature's output columns.

The same unit tables drive the fused NVLink kernels in ``torchrec_b200.parallel.p2p`` (pooled rows
written straight into the destination rank's output).
"""
from __future__ import annotations

import itertools
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..modules.embedding_configs import BaseEmbeddingConfig, DataType, PoolingType, data_type_to_dtype
from ..ops import jagged as J
from ..ops.tbe import OptimType, PoolingMode, TableBatchedEmbeddingBags, WeightDecayMode
from ..sparse.jagged_tensor import KeyedJaggedTensor
from .dist_data import KJTAllToAll, PooledEmbeddingsAllToAll, SequenceEmbeddingsAllToAll
from .sharding_plan import placement_rank
from .types import Awaitable, NoWait, ParameterSharding, QuantizedCommCodecs, ShardingEnv, ShardingType


@dataclass
class TableShard:
    table_idx: int
    name: str
    rank: int
    row_off: int
    col_off: int
    rows: int
    cols: int
    group: int = 0
    local_idx: int = -1  # index among the owner's tables of the group


@dataclass
class Unit:
    feature: int  # flat feature index in the module's embedding-name order
    shard: TableShard
    gidx: int = -1  # global unit index (rank-major)


@dataclass
class OptimizerSpec:
    """Sparse optimizer of one table (from ``apply_optimizer_in_backward`` tags or fused_params)."""

    optim: OptimType = OptimType.EXACT_SGD
    lr: float = 0.01
    eps: float = 1.0e-8
    beta1: float = 0.9
    beta2: float = 0.999
    weight_decay: float = 0.0
    weight_decay_mode: WeightDecayMode = WeightDecayMode.NONE
    max_gradient: float = 0.0
    momentum: float = 0.0
    stochastic_rounding: bool = True  # fp16 / bf16 tables only (fused_params["stochastic_rounding"], FBGEMM default True)

    def key(self) -> Tuple:
        return (self.optim, self.lr, self.eps, self.beta1, self.beta2, self.weight_decay, int(self.weight_decay_mode), self.max_gradient, self.momentum,
                bool(self.stochastic_rounding))


def shards_of(table_idx: int, cfg: BaseEmbeddingConfig, ps: ParameterSharding) -> List[TableShard]:
    out: List[TableShard] = []
    spec = ps.sharding_spec
    assert spec is not None, f"table {cfg.name}: model-parallel sharding needs a sharding_spec"
    for sm in spec.shards:  # type: ignore[attr-defined]
        out.append(TableShard(table_idx, cfg.name, placement_rank(sm.placement), sm.shard_offsets[0], sm.shard_offsets[1],
                              sm.shard_sizes[0], sm.shard_sizes[1]))
    return out


class _Group:
    """Tables that can share one table-batched kernel on a rank (same dtype / pooling / optimizer)."""

    def __init__(self, key: Tuple, pooling: PoolingMode, dtype: torch.dtype, opt: OptimizerSpec) -> None:
        self.key = key
        self.pooling = pooling
        self.dtype = dtype
        self.opt = opt
        self.local_shards: List[TableShard] = []  # this rank's shards, unit order
        self.unit_range: Tuple[int, int] = (0, 0)  # range in this rank's local unit list
        self.col_range: Tuple[int, int] = (0, 0)  # columns in the local output buffer
        self.tbe: Optional[TableBatchedEmbeddingBags] = None


class ShardedLookupEngine(nn.Module):
    """Input dist + lookup + output dist for the model-parallel tables of one embedding module."""

    def __init__(
        self,
        tables: List[BaseEmbeddingConfig],
        feature_names: List[str],  # flat (per table, per feature) input feature names
        feature_table: List[int],  # table index of each flat feature
        plan: Dict[str, ParameterSharding],
        env: ShardingEnv,
        device: torch.device,
        pooled: bool,
        is_weighted: bool,
        opt_specs: Dict[str, OptimizerSpec],
        output_dtype: torch.dtype = torch.float32,
        qcomm_codecs: Optional[QuantizedCommCodecs] = None,
    ) -> None:
        super().__init__()
        self._env = env
        self._pg = env.process_group
        self._W = env.world_size
        self._rank = env.rank
        self._device = device
        self._pooled = pooled
        self._is_weighted = is_weighted
        self._tables = tables
        self._feature_names = feature_names
        self._feature_table = feature_table
        self._output_dtype = output_dtype
        self._codecs = qcomm_codecs
        F = len(feature_names)

        # ---- shards, groups ------------------------------------------------------------------
        self._table_shards: Dict[int, List[TableShard]] = {}
        self._mp_tables: List[int] = []
        group_index: Dict[Tuple, int] = {}
        self._groups: List[_Group] = []
        self._post_mean_feature: List[bool] = [False] * F  # SUM in kernel, divide after the reduce
        for ti, cfg in enumerate(tables):
            ps = plan[cfg.name]
            if ps.sharding_type == ShardingType.DATA_PARALLEL.value:
                continue
            self._mp_tables.append(ti)
            shards = shards_of(ti, cfg, ps)
            self._table_shards[ti] = shards
            n_row_shards = len({(s.row_off, s.rows) for s in shards})
            pooling = PoolingMode.NONE
            if pooled:
                pt = getattr(cfg, "pooling", PoolingType.SUM)
                pooling = PoolingMode.MEAN if pt == PoolingType.MEAN else PoolingMode.SUM
                if pooling == PoolingMode.MEAN and n_row_shards > 1:
                    pooling = PoolingMode.SUM
                    for fi in range(F):
                        if feature_table[fi] == ti:
                            self._post_mean_feature[fi] = True
            dtype = torch.float32 if cfg.data_type == DataType.FP32 else data_type_to_dtype(cfg.data_type)
            opt = opt_specs.get(cfg.name, OptimizerSpec())
            # storage location: compute kernel of the plan (HBM | zero-copy host | host + HBM cache), reference embedding_types.py:75-95
            ck = getattr(ps, "compute_kernel", None) or "fused"
            # 1 zero-copy host rows, 2 host rows + HBM cache, 3 key-value virtual table in DRAM, 4 key-value virtual table on SSD
            loc = {"fused_uvm": 1, "fused_uvm_caching": 2, "key_value": 3, "dram_virtual_table": 3, "ssd_virtual_table": 4}.get(str(ck), 0) \
                if device.type == "cuda" or os.environ.get("TRB_UVM_ON_CPU") else 0
            clf = 0.2
            cp = getattr(ps, "cache_params", None)
            if cp is not None and getattr(cp, "load_factor", None):
                clf = float(cp.load_factor)
            kvp = getattr(ps, "key_value_params", None)
            kv_dir = getattr(kvp, "ssd_storage_directory", None) if kvp is not None else None
            # sequence (unpooled) kernels emit rows of ONE width per launch: column slices and whole tables of a collection (a column-wise
            # table next to a row-wise one) go to separate groups; ``lookup`` pads their rows to the widest unit of the job
            key = (str(dtype), int(pooling), opt.key(), loc, clf if loc >= 2 else 0, kv_dir if loc == 4 else None, None if pooled else shards[0].cols)
            if key not in group_index:
                group_index[key] = len(self._groups)
                self._groups.append(_Group(key, pooling, dtype, opt))
            for s in shards:
                s.group = group_index[key]

        # ---- unit order: rank-major, then group, then table (config order), shard, feature ----
        feats_of_table: Dict[int, List[int]] = {}
        for fi, ti in enumerate(feature_table):
            feats_of_table.setdefault(ti, []).append(fi)
        self._units: List[Unit] = []
        self._units_per_rank: List[int] = []
        self._dim_sum_per_rank: List[int] = []
        self._table_row_sharded: Dict[int, bool] = {
            ti: len({(s.row_off, s.rows) for s in self._table_shards[ti]}) > 1 for ti in self._mp_tables
        }
        # inside a group, tables that need no cross-rank reduction come first (lets the fused NVLink
        # path cover them with one launch writing straight into the destination output)
        ordered_tables = sorted(self._mp_tables, key=lambda ti: (self._table_row_sharded[ti], ti))
        for r in range(self._W):
            n0 = len(self._units)
            dsum = 0
            for gi, g in enumerate(self._groups):
                u0 = len(self._units) - n0
                c0 = dsum
                for ti in ordered_tables:
                    for s in self._table_shards[ti]:
                        if s.rank != r or s.group != gi:
                            continue
                        if r == self._rank:
                            s.local_idx = len(g.local_shards)
                            g.local_shards.append(s)
                        for fi in feats_of_table.get(ti, []):
                            self._units.append(Unit(fi, s, len(self._units)))
                            dsum += s.cols if pooled else 0
                if r == self._rank:
                    g.unit_range = (u0, len(self._units) - n0)
                    g.col_range = (c0, dsum)
            self._units_per_rank.append(len(self._units) - n0)
            self._dim_sum_per_rank.append(dsum)
        self._unit_start = list(itertools.accumulate([0] + self._units_per_rank))
        self._local_units = self._units[self._unit_start[self._rank] : self._unit_start[self._rank + 1]]
        self._mp_features: List[int] = sorted({u.feature for u in self._units})
        self._has_mp = len(self._units) > 0

        # ---- local kernels -------------------------------------------------------------------------
        self._tbes = nn.ModuleList()
        for gi, g in enumerate(self._groups):
            if not g.local_shards:
                self._tbes.append(nn.Identity())
                continue
            u0, u1 = g.unit_range
            units = self._local_units[u0:u1]
            loc, clf = g.key[3], g.key[4]
            extra: Dict[str, Any] = {}
            cls = TableBatchedEmbeddingBags
            if loc == 1:
                from ..ops.tbe import EmbeddingLocation

                extra["location"] = EmbeddingLocation.MANAGED
            elif loc == 2:
                from ..ops.uvm import UvmCachedEmbeddingBags

                cls = UvmCachedEmbeddingBags
                extra["cache_load_factor"] = clf
            elif loc in (3, 4):
                from ..ops.kv_tbe import KeyValueEmbeddingBags

                cls = KeyValueEmbeddingBags
                extra["cache_load_factor"] = clf
                extra["backend"] = "ssd" if loc == 4 else "dram"
                if loc == 4 and g.key[5]:
                    extra["ssd_storage_directory"] = os.path.join(g.key[5], f"rank{self._rank}_group{gi}")
                cap = int(os.environ.get("TRB_KV_STORE_ROWS", "0"))
                if cap:
                    extra["store_rows"] = [min(s.rows, cap) for s in g.local_shards]
            tbe = cls(
                embedding_specs=[(s.rows, s.cols) for s in g.local_shards],
                feature_table_map=[u.shard.local_idx for u in units],
                pooling_mode=g.pooling,
                weights_precision=g.dtype,
                output_dtype=output_dtype,
                optimizer=g.opt.optim,
                learning_rate=g.opt.lr, eps=g.opt.eps, beta1=g.opt.beta1, beta2=g.opt.beta2,
                weight_decay=g.opt.weight_decay, weight_decay_mode=g.opt.weight_decay_mode,
                max_gradient=g.opt.max_gradient, momentum=g.opt.momentum,
                device=device, table_names=[s.name for s in g.local_shards], **extra,
            )
            if cls is TableBatchedEmbeddingBags:
                tbe.stochastic_rounding = bool(g.opt.stochastic_rounding) and g.dtype in (torch.float16, torch.bfloat16)
            g.tbe = tbe
            self._tbes.append(tbe)

        self._build_routes()
        self._build_combine()

        # ---- portable transport modules ---------------------------------------------------------------
        self._p2p: Optional["_P2PState"] = None
        self._p2p_in: Optional["_P2PInputState"] = None
        self._p2p_checked = False
        self._kjt_a2a: Optional[KJTAllToAll] = None
        self._pooled_a2a: Optional[PooledEmbeddingsAllToAll] = None
        self._seq_a2a: Optional[SequenceEmbeddingsAllToAll] = None
        if self._has_mp and self._pg is not None and self._W > 1:
            self._kjt_a2a = KJTAllToAll(self._pg, self._units_per_rank)
            if pooled:
                self._pooled_a2a = PooledEmbeddingsAllToAll(self._pg, self._dim_sum_per_rank, device, codecs=qcomm_codecs)
            else:
                self._seq_a2a = SequenceEmbeddingsAllToAll(self._pg, self._units_per_rank, device, codecs=qcomm_codecs)

    # ---------------------------------------------------------------------------------------------------
    def _build_routes(self) -> None:
        """Host tables used to turn the local KJT into the routed (unit-ordered) KJT."""
        F = len(self._feature_names)
        units_by_feature: Dict[int, List[Unit]] = {}
        for u in self._units:
            units_by_feature.setdefault(u.feature, []).append(u)
        self._row_sharded = any(
            len({(u.shard.row_off, u.shard.rows) for u in us}) > 1 for us in units_by_feature.values()
        )
        # fast path: routed KJT is a pure key permutation (with repeats for column shards)
        self._perm_features = [self._mp_features.index(u.feature) for u in self._units]
        self._routed_keys = [self._feature_names[u.feature] for u in self._units]
        if not self._row_sharded:
            return
        # general path tables over the (feature, column-slice) pairs k and row shards j
        mp_pos = {f: i for i, f in enumerate(self._mp_features)}
        fc_count = [0] * len(self._mp_features)
        fc_base = [0] * len(self._mp_features)
        keyed_bounds: List[int] = []
        row_lo_flat: List[int] = []
        unit_flat: List[int] = []
        k = 0
        max_rows = max(self._tables[ti].num_embeddings for ti in self._mp_tables) + 1
        self._BIG = max_rows
        for f in self._mp_features:
            us = units_by_feature[f]
            cols = sorted({u.shard.col_off for u in us})
            fc_base[mp_pos[f]] = k
            fc_count[mp_pos[f]] = len(cols)
            for c in cols:
                rs = sorted([u for u in us if u.shard.col_off == c and u.shard.rows > 0], key=lambda u: u.shard.row_off)
                if not rs:
                    rs = [u for u in us if u.shard.col_off == c][:1]
                for u in rs:
                    keyed_bounds.append(k * self._BIG + u.shard.row_off)
                    row_lo_flat.append(u.shard.row_off)
                    unit_flat.append(u.gidx)
                k += 1
        dev = self._device
        self.register_buffer("_fc_count", torch.tensor(fc_count, dtype=torch.int64, device=dev), persistent=False)
        self.register_buffer("_fc_base", torch.tensor(fc_base, dtype=torch.int64, device=dev), persistent=False)
        self.register_buffer("_keyed_bounds", torch.tensor(keyed_bounds, dtype=torch.int64, device=dev), persistent=False)
        self.register_buffer("_row_lo_flat", torch.tensor(row_lo_flat, dtype=torch.int64, device=dev), persistent=False)
        self.register_buffer("_unit_flat", torch.tensor(unit_flat, dtype=torch.int64, device=dev), persistent=False)
        self._uniform_fc = len(set(fc_count)) == 1

    def _build_combine(self) -> None:
        """Map every received column (global unit order) to its output column."""
        if not self._pooled:
            return
        dims = [self._tables[ti].embedding_dim for ti in self._feature_table]
        self._out_base = list(itertools.accumulate([0] + dims))
        mp_cols: List[int] = []
        for u in self._units:
            base = self._out_base[u.feature] + u.shard.col_off
            mp_cols.extend(range(base, base + u.shard.cols))
        self._mp_dest_cols = mp_cols
        self._mp_is_perm = len(set(mp_cols)) == len(mp_cols)

    # ---- input dist ---------------------------------------------------------------------------------------
    def route(self, features: KeyedJaggedTensor) -> Tuple[KeyedJaggedTensor, Optional[torch.Tensor]]:
        """Local KJT (keys = model-parallel features in flat order) -> routed KJT (global unit
        order). Returns the unbucketize permutation for sequence outputs when ids were re-sorted."""
        if not self._row_sharded:
            if self._perm_features == list(range(len(features.keys()))):
                return features, None
            return features.permute(self._perm_features), None
        if features.values().is_cuda and self._uniform_fc and not features.variable_stride_per_key() and os.environ.get("TRB_ROUTE_EAGER") != "1":
            return self._route_cuda(features)
        B = features.stride()
        Fm = len(self._mp_features)
        lengths = features.lengths().to(torch.int64)
        values = features.values()
        n = values.numel()
        dev = values.device
        bag = torch.repeat_interleave(torch.arange(Fm * B, device=dev), lengths, output_size=n)
        f = torch.div(bag, B, rounding_mode="floor")
        b = bag - f * B
        rep = self._fc_count[f]
        if self._uniform_fc:
            total = n * int(self._fc_count[0]) if Fm else 0
        else:
            total = int(rep.sum())
        e = torch.repeat_interleave(torch.arange(n, device=dev), rep, output_size=total)
        rep_start = torch.cumsum(rep, 0) - rep
        c_local = torch.arange(total, device=dev) - rep_start[e]
        k = self._fc_base[f[e]] + c_local
        ids = values[e].to(torch.int64)
        key = k * self._BIG + ids.clamp(min=0, max=self._BIG - 1)
        pos = torch.searchsorted(self._keyed_bounds, key, right=True) - 1
        unit = self._unit_flat[pos]
        local_id = ids - self._row_lo_flat[pos]
        U = len(self._units)
        new_bag = unit * B + b[e]
        order = torch.argsort(new_bag, stable=True)
        new_lengths = torch.bincount(new_bag, minlength=U * B).to(features.lengths().dtype)
        new_values = local_id[order].to(values.dtype)
        w = features.weights_or_none()
        new_weights = None if w is None else w[e][order]
        unbucketize = torch.empty(total, dtype=torch.int64, device=dev)
        unbucketize[order] = torch.arange(total, device=dev)
        routed = KeyedJaggedTensor(keys=self._routed_keys, values=new_values, weights=new_weights, lengths=new_lengths, stride=B)
        return routed, unbucketize

    def _route_cuda(self, features: KeyedJaggedTensor) -> Tuple[KeyedJaggedTensor, Optional[torch.Tensor]]:
        """Row-sharded routing on the GPU for the portable (NCCL) transport: the same device-side pass as the NVLink plane's input dist
        (``csrc/kjt_route.cu``: bucketize by row range + rebase + permute into unit order + unbucketize permutation) with ONE local
        destination. No eager index arithmetic, no host sync: every id lands in exactly ``fc`` units (its row shard x the column slices),
        so the routed KJT holds ``n * fc`` ids."""
        import ctypes

        from ..ops import _lib

        dev = features.values().device
        tab = self.__dict__.get("_route_tables")
        if tab is None or tab["dev"] != dev:
            mp_pos = {f: i for i, f in enumerate(self._mp_features)}
            INT64_MAX = (1 << 63) - 1
            cols_of_feature: Dict[int, List[int]] = {}
            for u in self._units:
                cols_of_feature.setdefault(u.feature, [])
                if u.shard.col_off not in cols_of_feature[u.feature]:
                    cols_of_feature[u.feature].append(u.shard.col_off)
            for f in cols_of_feature:
                cols_of_feature[f].sort()
            rs = self._table_row_sharded
            mk32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)
            mk64 = lambda x: torch.tensor(x, dtype=torch.int64, device=dev)
            U = len(self._units)
            # first / last row shard of a table also take the ids below / above the table (forwarded as invalid ids), so every id is routed
            row_los = {ti: min(s.row_off for s in self._table_shards[ti] if s.rows > 0) if any(s.rows > 0 for s in self._table_shards[ti]) else 0 for ti in self._mp_tables}
            row_his = {ti: max(s.row_off + s.rows for s in self._table_shards[ti]) for ti in self._mp_tables}
            edge = [((1 if u.shard.row_off == row_los[u.shard.table_idx] and u.shard.rows > 0 else 0) | (2 if u.shard.row_off + u.shard.rows == row_his[u.shard.table_idx] and u.shard.rows > 0 else 0))
                    if rs[u.shard.table_idx] else 0 for u in self._units]
            tab = self.__dict__["_route_tables"] = {
                "dev": dev, "U": U, "fc": int(self._fc_count[0]) if len(self._mp_features) else 1,
                "key": mk32([mp_pos[u.feature] for u in self._units]),
                "lo": mk64([u.shard.row_off if rs[u.shard.table_idx] else 0 for u in self._units]),
                "hi": mk64([u.shard.row_off + u.shard.rows if rs[u.shard.table_idx] else INT64_MAX for u in self._units]),
                "zero": mk32([0] * U), "slot": mk32(list(range(U))), "cslice": mk32([cols_of_feature[u.feature].index(u.shard.col_off) for u in self._units]),
                "ustart": mk32([0, U]), "overflow": torch.zeros(1, dtype=torch.int32, device=dev), "edge": mk32(edge)}
        B = features.stride()
        values = features.values()
        n, fc, U = values.numel(), tab["fc"], tab["U"]
        total = n * fc
        L = _lib.lib()
        L.trb_kjt_route_workspace_bytes.restype = ctypes.c_int64
        ws = torch.empty(int(L.trb_kjt_route_workspace_bytes(U, B)), dtype=torch.uint8, device=dev)
        len32 = torch.empty(U * B + 1, dtype=torch.int32, device=dev)
        off32 = torch.empty(U * B + 1, dtype=torch.int32, device=dev)
        new_values = torch.empty(total, dtype=values.dtype, device=dev)
        w = features.weights_or_none()
        w32 = w.float().contiguous() if w is not None else None
        new_w = torch.empty(total, dtype=torch.float32, device=dev) if w is not None else None
        unb = torch.empty(total, dtype=torch.int64, device=dev)
        in_off = features.offsets()
        code = L.trb_kjt_route_ex(
            _lib.ptr(in_off), int(in_off.dtype == torch.int64), _lib.ptr(values), int(values.dtype == torch.int64), _lib.ptr(w32), B, _lib.ptr(tab["key"]),
            _lib.ptr(tab["lo"]), _lib.ptr(tab["hi"]), _lib.ptr(tab["zero"]), _lib.ptr(tab["slot"]), _lib.ptr(tab["cslice"]), _lib.ptr(tab["ustart"]), U, 1,
            _lib.ptr_array([off32.data_ptr()]), 0, _lib.ptr_array([new_values.data_ptr()]), int(values.dtype == torch.int64),
            _lib.ptr_array([new_w.data_ptr()]) if new_w is not None else ctypes.c_void_p(0), ctypes.c_int64(max(total, 1)), _lib.ptr(unb), fc, ctypes.c_void_p(0),
            ctypes.c_void_p(0), ctypes.c_void_p(0), 1, _lib.ptr(tab["edge"]), _lib.ptr(len32), _lib.ptr(tab["overflow"]), _lib.ptr(ws), ctypes.c_int64(ws.numel()),
            int(max(1, n // max(len(self._mp_features) * B, 1))), _lib.stream_ptr(dev))
        _lib.check(code, "trb_kjt_route_ex")
        routed = KeyedJaggedTensor(keys=self._routed_keys, values=new_values, weights=None if new_w is None else new_w.to(w.dtype),
                                   lengths=len32[: U * B].to(features.lengths().dtype), stride=B)
        return routed, unb

    def input_dist_routed(self, routed: KeyedJaggedTensor) -> Awaitable[Awaitable[KeyedJaggedTensor]]:
        """Input dist of a KJT that is already in global unit order (keys = unit features)."""
        if self._kjt_a2a is None:
            return NoWait(NoWait(routed))
        return self._kjt_a2a(routed)

    def input_dist(self, features: KeyedJaggedTensor) -> Tuple[Awaitable[Awaitable[KeyedJaggedTensor]], Optional[torch.Tensor]]:
        routed, unbucketize = self.route(features)
        if self._kjt_a2a is None:
            return NoWait(NoWait(routed)), unbucketize
        return self._kjt_a2a(routed), unbucketize

    def _uniform_batch(self, B: int) -> bool:
        """All ranks feed the same local batch size (checked once per size; required by the fixed-slot NVLink buffers)."""
        if self._W == 1 or self._pg is None:
            return True
        cache = self.__dict__.setdefault("_uniform_batch_cache", {})
        if B not in cache:
            t = torch.tensor([B, -B], device=self._device, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self._pg)
            cache[B] = int(t[0].item()) == B and int(-t[1].item()) == B
        return cache[B]

    # ---- lookup -------------------------------------------------------------------------------------------
    def lookup(self, dist_features: KeyedJaggedTensor) -> torch.Tensor:
        """Run the table-batched kernels over this rank's units. Pooled: ``[B_global, D_local]``;
        sequence: ``[n_values, D]`` (received order)."""
        self._run_lookup_hooks(dist_features)
        Bg = dist_features.stride()
        values = dist_features.values()
        offsets = dist_features.offsets()
        weights = dist_features.weights_or_none() if self._is_weighted else None
        outs: List[torch.Tensor] = []
        for g in self._groups:
            if g.tbe is None:
                continue
            u0, u1 = g.unit_range
            window = offsets[u0 * Bg : u1 * Bg + 1]
            if self._pooled:
                outs.append(g.tbe(values, window, weights, batch_size=Bg))
            else:
                outs.append(g.tbe(values, window, None, batch_size=Bg))
        if not self._pooled:
            Wd = self.__dict__.get("_seq_width")
            if Wd is None:
                Wd = self.__dict__["_seq_width"] = max((u.shard.cols for u in self._units), default=self._tables[0].embedding_dim if self._tables else 0)
            outs = [torch.nn.functional.pad(o, (0, Wd - o.shape[1])) if o.shape[1] < Wd else o for o in outs]
        if not outs:
            D = 0 if self._pooled else self.__dict__["_seq_width"]
            # a rank without shards of this module still has to run the BACKWARD collectives of the output dist with its peers:
            # the empty result must be part of the autograd graph (found by the tower test: TW table on one rank only -> hang)
            return torch.zeros(Bg if self._pooled else 0, D, device=values.device, dtype=self._output_dtype, requires_grad=torch.is_grad_enabled())
        if len(outs) == 1:
            return outs[0]
        if self._pooled:
            return torch.cat(outs, dim=1)
        # sequence groups own disjoint value windows: trim each to its window and concatenate
        trimmed = []
        for g, o in zip([g for g in self._groups if g.tbe is not None], outs):
            u0, u1 = g.unit_range
            lo, hi = int(offsets[u0 * Bg]), int(offsets[u1 * Bg])
            trimmed.append(o[: hi - lo])
        return torch.cat(trimmed, 0)

    @torch.no_grad()
    def prefetch(self, dist_features: KeyedJaggedTensor) -> None:
        """Make the rows of the next batch resident in the HBM caches (UVM_CACHING groups); no-op otherwise.
        Called by PrefetchTrainPipelineSparseDist on its prefetch stream (reference embedding_lookup.py:714-767)."""
        Bg = dist_features.stride()
        offsets = dist_features.offsets()
        for g in self._groups:
            if g.tbe is None or not getattr(g.tbe, "is_cached", False):
                continue
            u0, u1 = g.unit_range
            g.tbe.prefetch(dist_features.values(), offsets[u0 * Bg : u1 * Bg + 1], Bg)

    # ---- output dist (pooled) ---------------------------------------------------------------------------------
    def output_dist(self, local_embs: torch.Tensor, batch_size_per_rank: Optional[List[int]] = None) -> Awaitable[torch.Tensor]:
        if self._pooled_a2a is None:
            return NoWait(local_embs)
        return self._pooled_a2a(local_embs, batch_size_per_rank)

    def combine(self, mp_embs: Optional[torch.Tensor], dp_embs: Optional[torch.Tensor], dp_cols: List[int], total_cols: int,
                mean_divisor: Optional[torch.Tensor]) -> torch.Tensor:
        """Place received unit columns (and data-parallel columns) into the module's output layout."""
        parts, cols = [], []
        if mp_embs is not None and mp_embs.shape[1] > 0:
            parts.append(mp_embs)
            cols.extend(self._mp_dest_cols)
        if dp_embs is not None and dp_embs.shape[1] > 0:
            parts.append(dp_embs.to(parts[0].dtype) if parts else dp_embs)
            cols.extend(dp_cols)
        x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        if cols == list(range(total_cols)):
            out = x
        else:
            key = (tuple(cols), total_cols, x.device)
            idx = self._combine_cache.get(key) if hasattr(self, "_combine_cache") else None
            if idx is None:
                if not hasattr(self, "_combine_cache"):
                    self._combine_cache: Dict[Any, Tuple[bool, torch.Tensor]] = {}
                is_perm = len(set(cols)) == len(cols) and len(cols) == total_cols
                if is_perm:
                    inv = [0] * total_cols
                    for src, dst in enumerate(cols):
                        inv[dst] = src
                    with torch.inference_mode(False):
                        idx = (True, torch.tensor(inv, dtype=torch.int64, device=x.device))
                else:
                    with torch.inference_mode(False):
                        idx = (False, torch.tensor(cols, dtype=torch.int64, device=x.device))
                self._combine_cache[key] = idx
            if idx[0]:
                out = x.index_select(1, idx[1])
            else:
                out = torch.zeros(x.shape[0], total_cols, dtype=x.dtype, device=x.device).index_add(1, idx[1], x)
        if mean_divisor is not None:
            out = out * mean_divisor
        return out

    # ---- misc -----------------------------------------------------------------------------------------------
    def register_lookup_hook(self, fn):
        """``fn(engine, dist_features)`` runs before every lookup over this rank's units (model delta tracker).
        Returns a callable that removes the hook."""
        hooks = self.__dict__.setdefault("_lookup_hooks", [])
        hooks.append(fn)
        return lambda: hooks.remove(fn) if fn in hooks else None

    def _run_lookup_hooks(self, dist_features: KeyedJaggedTensor) -> None:
        for fn in self.__dict__.get("_lookup_hooks", ()):
            fn(self, dist_features)

    @property
    def local_units(self) -> List[Unit]:
        return self._local_units

    @property
    def units(self) -> List[Unit]:
        return self._units

    @property
    def units_per_rank(self) -> List[int]:
        return self._units_per_rank

    @property
    def dim_sum_per_rank(self) -> List[int]:
        return self._dim_sum_per_rank

    @property
    def groups(self) -> List[_Group]:
        return self._groups

    @property
    def mp_features(self) -> List[int]:
        return self._mp_features

    def local_shard_views(self) -> List[Tuple[TableShard, torch.Tensor, Dict[str, torch.Tensor], TableBatchedEmbeddingBags]]:
        """(shard, weight view [rows, cols], optimizer state views, owning kernel) for every local shard."""
        res = []
        for g in self._groups:
            if g.tbe is None:
                continue
            ws = g.tbe.split_embedding_weights()
            sts = g.tbe.split_optimizer_states()
            for s in g.local_shards:
                res.append((s, ws[s.local_idx], sts[s.local_idx], g.tbe))
        return res


    @torch.no_grad()
    def reset_rows(self, table: str, global_rows: torch.Tensor) -> int:
        """Re-initialise the given (global) rows of ``table`` wherever this rank holds them — weights back to the
        table's init range, optimizer state to zero. Used by managed-collision eviction and ITEP pruning."""
        n_reset = 0
        for shard, w, st, _tbe in self.local_shard_views():
            if shard.name != table or shard.rows == 0:
                continue
            cfg = self._tables[shard.table_idx]
            ids = global_rows.to(w.device).long()
            ids = ids[(ids >= shard.row_off) & (ids < shard.row_off + shard.rows)] - shard.row_off
            if ids.numel() == 0:
                continue
            w[ids] = torch.empty(ids.numel(), w.shape[1], device=w.device, dtype=torch.float32).uniform_(cfg.get_weight_init_min(), cfg.get_weight_init_max()).to(w.dtype)
            for v in st.values():
                v[ids] = 0
            n_reset += int(ids.numel())
        return n_reset

    # ---- fused NVLink path (single NVLink domain) -----------------------------------------------------------------
    def fused_available(self, batch_size_per_rank: Optional[List[int]]) -> bool:
        """The fused lookup + output-dist kernels can serve this batch (CUDA, one host, even batch)."""
        if not self._pooled or not self._has_mp or self._device.type != "cuda":
            return False
        if self._W == 1 and os.environ.get("TRB_PLANE_SINGLE", "1") == "0":
            return False
        if any(getattr(g.tbe, "is_cached", False) for g in self._groups):
            return False  # cached tables translate ids -> cache slots first: portable path
        if batch_size_per_rank is not None and len(set(batch_size_per_rank)) != 1:
            return False
        if getattr(self._env, "loopback_group", None) is not None or self._W == 1:
            return True
        if not self._p2p_checked:
            from .p2p import PeerGroup

            self._p2p_ok = PeerGroup.supported(self._pg, self._device)
            self._p2p_checked = True
        return self._p2p_ok

    # ---- NVLink sparse plane (see sparse_plane.py) ------------------------------------------------------------------------
    def _peer_group(self):
        lb = getattr(self._env, "loopback_group", None)
        if lb is not None:
            return lb.view(self._rank)
        if self._W == 1:
            from .sparse_plane import SingleRankGroup

            return SingleRankGroup(self._device)
        from .p2p import PeerGroup

        return PeerGroup.get(self._pg, self._device)

    def _region_capacity(self, features: KeyedJaggedTensor, key_of_feature: List[int]) -> int:
        """Ids one source may send to one destination per batch: measured on the first batch (max over destinations and ranks),
        exact when every feature has a fixed bag length, with head-room (TRB_ID_CAPACITY_SLACK, default 1.5) otherwise."""
        env_cap = int(os.environ.get("TRB_MAX_IDS_PER_RANK", "0"))
        if env_cap:
            return env_cap
        B = features.stride()
        lengths = features.lengths().view(-1, B)
        values = features.values()
        offs = features.offsets()
        per_dest = torch.zeros(self._W, dtype=torch.int64, device=values.device)
        lpk = lengths.sum(1)
        fixed = bool((lengths == lengths[:, :1]).all().item())
        for u in self._units:
            k = key_of_feature[u.feature]
            if self._table_row_sharded[u.shard.table_idx]:
                seg = values[int(offs[k * B]) : int(offs[(k + 1) * B])]
                per_dest[u.shard.rank] += ((seg >= u.shard.row_off) & (seg < u.shard.row_off + u.shard.rows)).sum()
                fixed = False
            else:
                per_dest[u.shard.rank] += lpk[k]
        n = per_dest.max().reshape(1)
        flag = torch.tensor([0 if fixed else 1], device=n.device, dtype=torch.int64)
        if self._pg is not None and self._W > 1 and getattr(self._env, "loopback_group", None) is None:
            dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self._pg)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self._pg)
        if int(flag.item()) == 0:
            return max(int(n.item()), 32)
        slack = float(os.environ.get("TRB_ID_CAPACITY_SLACK", "1.5"))
        return int(int(n.item()) * slack) + 1024

    def plane_for(self, B_local: int, total_cols: int, features: Optional[KeyedJaggedTensor] = None, key_of_feature: Optional[List[int]] = None,
                  capacity: Optional[int] = None):
        """The sparse plane serving batches of ``B_local`` samples per rank (built on first use; collective)."""
        from .sparse_plane import SparsePlane

        planes = self.__dict__.setdefault("_planes", {})
        has_input = features is not None
        key = (B_local, total_cols, has_input)
        pl = planes.get(key)
        if pl is None:
            if has_input:
                kof = list(key_of_feature) if key_of_feature is not None else None
                if kof is None:
                    mp_pos = {f: i for i, f in enumerate(self._mp_features)}
                    kof = [mp_pos.get(f, 0) for f in range(len(self._feature_names))]
                cap = int(capacity) if capacity else self._region_capacity(features, kof)
                pl = SparsePlane(self, self._peer_group(), B_local, total_cols, cap, features.values().dtype, self._is_weighted and features.weights_or_none() is not None, kof)
            else:
                pl = SparsePlane(self, self._peer_group(), B_local, total_cols, 0, torch.int64, self._is_weighted, None)
            planes[key] = pl
        return pl

    def plane_input_dist(self, features: KeyedJaggedTensor, key_of_feature: Optional[List[int]], total_cols: int, capacity: Optional[int] = None,
                         training: Optional[bool] = None):
        """NVLink input dist: ONE device-side pass (bucketize + feature permute + peer write, csrc/kjt_route.cu) straight from the
        batch's KJT into the owners' receive regions. No host sync, no size exchange. Returns a ``RoutedIds`` handle."""
        B = features.stride()
        pl = self.plane_for(B, total_cols, features, key_of_feature, capacity)
        n_bags = max(1, len(features.keys()) * B)
        return pl.push_input(features.offsets(), features.values(), features.weights_or_none() if self._is_weighted else None,
                             avg_len_hint=max(1, features.values().numel() // n_bags), training=training)

    def fused_lookup_dist(self, dist_features, B_local: int, total_cols: int, grad_scale: float, dp: Optional[Tuple] = None) -> torch.Tensor:
        """Lookup + pooled output dist in one pass: pooled rows are written straight into the owning
        rank's ``[B_local, total_cols]`` output over NVLink (row-sharded tables via staging slabs
        reduced at the destination). ``dist_features``: a ``RoutedIds`` handle of the NVLink input dist or a distributed KJT.
        Returns the local output (final column layout)."""
        from .sparse_plane import RoutedIds

        routed = isinstance(dist_features, RoutedIds)
        if self.__dict__.get("_lookup_hooks"):
            self._run_lookup_hooks(dist_features.to_kjt() if routed else dist_features)
        pl = dist_features.plane if routed else self.plane_for(B_local, total_cols)
        anchor = None
        for g in self._groups:
            if g.tbe is not None:
                anchor = g.tbe._dummy
                break
        if anchor is None:
            anchor = pl.dummy
        weights = None if routed else (dist_features.weights_or_none() if self._is_weighted else None)
        # dp = (dense TBE of the replicated tables, its meta with output columns, local KJT of their features, process group):
        # their rows are looked up locally and written into the same output buffer; the dense gradient is all-reduced in backward
        dp_w = dp[0].weights if dp is not None else None
        return _PlaneLookupFn.apply(anchor, self, pl, dist_features, weights, grad_scale, dp_w, dp)


class _PlaneLookupFn(torch.autograd.Function):
    """Forward: table-batched lookup writing pooled rows into the destination ranks' outputs over NVLink + device barrier
    (+ staging reduce). Backward: gradient column blocks pushed into the owners' inboxes, barrier, exact fused
    backward + optimizer over local memory. Both halves replay as CUDA graphs when the ids came through the plane."""

    @staticmethod
    def forward(ctx, anchor, eng: ShardedLookupEngine, pl, ids, weights, grad_scale: float, dp_weights=None, dp=None):
        from ..ops import tbe as T

        for tbe in eng._tbes:  # FULLY_SHARDED 2D strategy: full weights only around the kernels
            if tbe.__dict__.get("_fs") is not None:
                tbe._fs.before_forward()
        training = any(ctx.needs_input_grad)  # grad mode is off inside Function.forward: this is the caller's view
        out, out_slot = pl.forward(ids, training)
        if dp is not None:  # replicated tables: local lookup straight into their columns (nobody else writes them)
            dp_tbe, dp_meta, dp_kjt, _pg = dp
            dp_psw = dp_kjt.weights_or_none() if eng._is_weighted else None
            T.pooled_forward(dp_meta, dp_weights.detach(), dp_kjt.values(), dp_kjt.offsets(), dp_psw, dp_kjt.stride(), dp_tbe.pooling_mode == T.PoolingMode.MEAN,
                             pl.wire_dtype, out=out)
        ctx.eng, ctx.pl, ctx.ids, ctx.grad_scale, ctx.dp = eng, pl, ids, grad_scale, dp
        ctx.has_weights = weights is not None
        if training:
            for tbe in eng._tbes:
                if tbe.__dict__.get("_fs") is not None:
                    tbe._fs.after_forward()
        return out

    @staticmethod
    def backward(ctx, grad):
        from ..ops import tbe as T

        eng, pl, ids = ctx.eng, ctx.pl, ctx.ids
        for tbe in eng._tbes:
            if tbe.__dict__.get("_fs") is not None:
                tbe._fs.before_backward()
        if grad.stride(1) != 1:
            grad = grad.contiguous()
        want_psw = ctx.has_weights and ctx.needs_input_grad[4]
        if want_psw:
            n = ids.values().numel()
            if pl.psw_grad_buf is None or pl.psw_grad_buf.numel() < n:
                pl.psw_grad_buf = torch.zeros(n, dtype=torch.float32, device=grad.device)
        pl.backward(ids, grad, ctx.grad_scale, want_psw)
        gpsw = None
        if want_psw:
            gpsw = pl.psw_grad_buf[: ids.values().numel()].to(ids.weights().dtype)
        g_dp = None
        if ctx.dp is not None and ctx.needs_input_grad[6]:
            dp_tbe, dp_meta, dp_kjt, pg = ctx.dp
            dp_psw = dp_kjt.weights_or_none() if eng._is_weighted else None
            gw = torch.zeros(dp_tbe.weights.numel(), dtype=torch.float32, device=grad.device)
            T.fused_backward(dp_meta, dp_tbe.weights.detach(), gw, None, dp_tbe.hyper_dev, dp_tbe.hyper_host, 8, 0, dp_kjt.values(), dp_kjt.offsets(), dp_psw,
                             dp_kjt.stride(), dp_tbe.pooling_mode == T.PoolingMode.MEAN, grad=grad)
            if pg is not None and eng._W > 1:
                dist.all_reduce(gw, group=pg)
                gw.div_(eng._W)
            g_dp = gw.to(dp_tbe.weights.dtype)
        return torch.zeros(1, dtype=torch.float32, device=grad.device), None, None, None, gpsw, None, g_dp, None
