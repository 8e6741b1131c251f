"""ShardedEmbeddingCollection + sharder: unpooled (sequence) embeddings
(reference torchrec/distributed/embedding.py:421-1832).

Same lookup-unit engine as the pooled path (``engine.py``): ids are routed to the owning units (with
optional per-feature dedup before the all-to-all), each rank gathers rows for the global batch and a
sequence all-to-all returns ``[sum L, D]`` rows to the ranks that own the samples; row-wise shards are
merged back into the original id order with the unbucketize permutation and column-wise shards are
concatenated along the embedding dim.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, Iterator, List, Mapping, Optional, Set, Tuple, Type

import torch
import torch.distributed as dist
from torch import nn

from ..modules.embedding_configs import DataType, EmbeddingConfig, data_type_to_dtype
from ..modules.embedding_modules import EmbeddingCollection, EmbeddingCollectionInterface, get_embedding_names_by_table
from ..ops import jagged as J
from ..optim.fused import FusedOptimizerModule
from ..optim.keyed import CombinedOptimizer, KeyedOptimizer
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from ..streamable import Multistreamable
from .embedding_types import BaseEmbeddingSharder, KJTList, ShardedEmbeddingModule
from .embeddingbag import EmbeddingFusedOptimizer, _local_pieces_with_offsets, _sharded_tensor_from_local, _sharded_view, _TableParam, optimizer_spec_from
from .engine import OptimizerSpec, ShardedLookupEngine, TableShard
from .types import Awaitable, CommOp, LazyAwaitable, NoWait, ParameterSharding, QuantizedCommCodecs, ShardedModule, ShardingEnv, ShardingType

_EC_INDEX_DEDUP: bool = False


def set_ec_index_dedup(val: bool) -> None:
    """Deduplicate ids per feature before the input all-to-all (reference embedding.py:165)."""
    global _EC_INDEX_DEDUP
    _EC_INDEX_DEDUP = val


def get_ec_index_dedup() -> bool:
    return _EC_INDEX_DEDUP


class EmbeddingCollectionContext(Multistreamable):
    def __init__(self) -> None:
        self.features: Optional[KeyedJaggedTensor] = None  # local (permuted) model-parallel features
        self.routed: Optional[KeyedJaggedTensor] = None
        self.unbucketize: Optional[torch.Tensor] = None
        self.reverse_indices: Optional[torch.Tensor] = None
        self.lookup_lpk: List[int] = []
        self.input_splits: List[int] = []
        self.output_splits: List[int] = []
        self.dp_features: Optional[KeyedJaggedTensor] = None
        self.sharding_contexts: List[Any] = []

    def record_stream(self, stream: torch.Stream) -> None:
        for k in (self.features, self.routed, self.dp_features):
            if k is not None:
                k.record_stream(stream)
        for t in (self.unbucketize, self.reverse_indices):
            if t is not None and t.is_cuda:
                t.record_stream(stream)


class EmbeddingCollectionAwaitable(LazyAwaitable[Dict[str, JaggedTensor]]):
    def __init__(self, finish) -> None:
        super().__init__()
        self._finish = finish

    def _wait_impl(self) -> Dict[str, JaggedTensor]:
        return self._finish()


class _SeqInputDist(Awaitable[Awaitable[KJTList]]):
    def __init__(self, inner, ctx: EmbeddingCollectionContext, routed: KeyedJaggedTensor, splits: List[int], W: int) -> None:
        super().__init__()
        self._inner, self._ctx, self._routed, self._splits, self._W = inner, ctx, routed, splits, W

    def _wait_impl(self) -> Awaitable[KJTList]:
        tensors_aw = self._inner.wait()
        ctx = self._ctx
        # values sent to / received from every rank (rows of the later sequence all-to-all)
        if hasattr(tensors_aw, "_input_splits"):
            labels = self._routed.dist_labels()
            vi = labels.index("values")
            ctx.output_splits = list(tensors_aw._input_splits[vi])  # what comes back == what was sent
            ctx.input_splits = list(tensors_aw._output_splits[vi])  # what is sent back == what was received
        return _SeqInputDistTensors(tensors_aw)


class _SeqInputDistTensors(Awaitable[KJTList]):
    def __init__(self, inner) -> None:
        super().__init__()
        self._inner = inner

    def _wait_impl(self) -> KJTList:
        return KJTList([self._inner.wait()])


class ShardedEmbeddingCollection(ShardedEmbeddingModule[KJTList, List[torch.Tensor], Dict[str, JaggedTensor], EmbeddingCollectionContext], FusedOptimizerModule):
    """Sharded ``EmbeddingCollection``: KJT -> Dict[embedding name, JaggedTensor]."""

    def __init__(self, module: EmbeddingCollectionInterface, table_name_to_parameter_sharding: Dict[str, ParameterSharding], env: ShardingEnv,
                 fused_params: Optional[Dict[str, Any]] = None, device: Optional[torch.device] = None,
                 qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None, use_index_dedup: bool = False, module_fqn: Optional[str] = None) -> None:
        super().__init__(qcomm_codecs_registry=qcomm_codecs_registry)
        self._env = env
        self._pg = env.process_group
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        self._embedding_configs: List[EmbeddingConfig] = module.embedding_configs()
        self._plan = table_name_to_parameter_sharding
        self._need_indices = module.need_indices()
        self._embedding_dim = module.embedding_dim()
        self._use_index_dedup = use_index_dedup or get_ec_index_dedup()
        tables = self._embedding_configs
        for t in tables:
            if self._plan[t.name].sharding_type == ShardingType.DATA_PARALLEL.value:
                raise ValueError("EmbeddingCollection tables use model-parallel sharding types (table/row/column wise); "
                                 f"data_parallel was requested for {t.name}")
        self._embedding_names: List[str] = [n for names in get_embedding_names_by_table(tables) for n in names]
        self._feature_names: List[str] = [f for t in tables for f in t.feature_names]
        self._feature_table: List[int] = [ti for ti, t in enumerate(tables) for _ in t.feature_names]
        src_params = dict(module.named_parameters()) if isinstance(module, nn.Module) else {}
        opt_specs: Dict[str, OptimizerSpec] = {}
        for t in tables:
            spec, _ = optimizer_spec_from(src_params.get(f"embeddings.{t.name}.weight"), fused_params)
            opt_specs[t.name] = spec
        if getattr(module, "_trb_opt_specs", None):  # re-sharding keeps the optimizers of the module being replaced
            opt_specs.update(module._trb_opt_specs)
        self._opt_specs = opt_specs
        codecs = qcomm_codecs_registry.get(CommOp.SEQUENCE_EMBEDDINGS_ALL_TO_ALL.name) if qcomm_codecs_registry else None
        self._engine = ShardedLookupEngine(tables=tables, feature_names=self._feature_names, feature_table=self._feature_table, plan=self._plan,
                                           env=env, device=self._device, pooled=False, is_weighted=False, opt_specs=opt_specs, qcomm_codecs=codecs)
        self._init_from(module)
        self.embeddings = nn.ModuleDict()
        self._table_params: Dict[str, nn.Parameter] = {}
        by_table = self._local_shards_by_table()
        for cfg in tables:
            shards = by_table.get(cfg.name, [])
            if shards:
                p = nn.Parameter(shards[0][1], requires_grad=False)
                p._in_backward_optimizers = [None]  # type: ignore[attr-defined]
                self.embeddings[cfg.name] = _TableParam(p)
                self._table_params[cfg.name] = p
            else:
                self.embeddings[cfg.name] = _TableParam(None)
        # unit bookkeeping: for every feature the units (column slices in order) that make its rows
        self._units_of_feature: Dict[int, List[int]] = {}
        for u in self._engine.units:
            self._units_of_feature.setdefault(u.feature, []).append(u.gidx)
        self._features_order: Optional[List[int]] = None
        self._has_features_permute = False
        self._optim: Optional[KeyedOptimizer] = None
        # shim so that EmbeddingFusedOptimizer can enumerate state the same way as for the pooled module
        self._embedding_bag_configs = tables

    # ---- helpers shared with the pooled module ----------------------------------------------------------------
    @torch.no_grad()
    def _init_from(self, module: nn.Module) -> None:
        if self._device.type == "meta":
            return
        tables = {t.name: t for t in self._embedding_configs}
        src = {}
        if hasattr(module, "embeddings"):
            for name, emb in module.embeddings.items():
                w = getattr(emb, "weight", None)
                if w is not None and w.device.type != "meta":
                    src[name] = w
        for shard, wview, _st, _tbe in self._engine.local_shard_views():
            cfg = tables[shard.name]
            if shard.name in src:
                wview.copy_(src[shard.name][shard.row_off : shard.row_off + shard.rows, shard.col_off : shard.col_off + shard.cols])
            elif wview.numel() > 0:
                wview.copy_(torch.empty(wview.shape, dtype=torch.float32, device=wview.device).uniform_(cfg.get_weight_init_min(), cfg.get_weight_init_max()))

    @property
    def engine(self):
        return self._engine

    def prefetch(self, ctx, dist_input) -> None:
        """Stage the rows of an already-distributed batch into the HBM caches of UVM_CACHING tables (prefetch pipeline)."""
        if self._engine is not None and len(dist_input) > 0:
            self._engine.prefetch(dist_input[0])

    def reset_rows(self, table: str, global_rows: torch.Tensor) -> int:
        """Re-initialise rows of a sharded table (managed-collision eviction / ITEP); see engine.reset_rows."""
        return self._engine.reset_rows(table, global_rows) if self._engine is not None else 0

    def _local_shards_by_table(self):
        res: Dict[str, List] = {}
        for item in self._engine.local_shard_views():
            res.setdefault(item[0].name, []).append(item)
        return res

    def _global_shards(self, name: str):
        from .sharding_plan import placement_rank

        spec = self._plan[name].sharding_spec
        return [(list(s.shard_offsets), list(s.shard_sizes), placement_rank(s.placement)) for s in spec.shards]  # type: ignore[union-attr]

    def _table_state_tensors(self):
        from .comm import get_local_size

        res = {}
        local_size = get_local_size(self._env.world_size)
        for cfg in self._embedding_configs:
            shards = self._local_shards_by_table().get(cfg.name)
            if not shards or cfg.name not in self._table_params:
                continue
            g = self._global_shards(cfg.name)
            st = {}
            for n in shards[0][2].keys():
                first = shards[0][2][n]
                if first.dim() == 1:
                    cols = sorted({o[1] for o, _, _ in g})
                    local = [(s_st[n], [s.row_off + cols.index(s.col_off) * cfg.num_embeddings], [s.rows]) for s, _, s_st, _ in shards]
                    gl = [([o[0] + cols.index(o[1]) * cfg.num_embeddings], [sz[0]], r) for o, sz, r in g]
                    size = [len(cols) * cfg.num_embeddings]
                else:
                    local = [(s_st[n], [s.row_off, s.col_off], [s.rows, s.cols]) for s, _, s_st, _ in shards]
                    gl, size = g, [cfg.num_embeddings, cfg.embedding_dim]
                st[n] = _sharded_view(local, gl, size, torch.float32, self._plan[cfg.name].sharding_type, self._env, self._pg, self._device.type, local_size)
            res[cfg.name] = (self._table_params[cfg.name], st)
        return res

    # ---- ShardedModule contract -----------------------------------------------------------------------------------
    def create_context(self) -> EmbeddingCollectionContext:
        return EmbeddingCollectionContext()

    def input_dist(self, ctx: EmbeddingCollectionContext, features: KeyedJaggedTensor) -> Awaitable[Awaitable[KJTList]]:
        if self._features_order is None:
            pos = {k: i for i, k in enumerate(features.keys())}
            self._features_order = [pos[f] for f in self._feature_names]
            self._has_features_permute = self._features_order != list(range(len(features.keys())))
        with torch.no_grad():
            if self._has_features_permute:
                features = features.permute(self._features_order)
            ctx.features = features
            lookup_features = features
            if self._use_index_dedup:
                lookup_features, ctx.reverse_indices = self._dedup(features)
            ctx.lookup_lpk = lookup_features.length_per_key()
            routed, unbucketize = self._engine.route(lookup_features)
            ctx.routed = routed
            ctx.unbucketize = unbucketize
            if self._engine._kjt_a2a is None:
                n = routed.values().numel()
                ctx.input_splits, ctx.output_splits = [n], [n]
                return NoWait(NoWait(KJTList([routed])))
            return _SeqInputDist(self._engine._kjt_a2a(routed), ctx, routed, self._engine.units_per_rank, self._env.world_size)

    def _dedup(self, features: KeyedJaggedTensor) -> Tuple[KeyedJaggedTensor, torch.Tensor]:
        """Per-feature unique ids; bag structure collapses to one bag per feature holding the unique ids."""
        F = len(features.keys())
        B = features.stride()
        lpk = features.length_per_key()
        values = features.values()
        uniq_vals, inverse, new_lengths = [], [], []
        off = 0
        base = 0
        for f in range(F):
            v = values[off : off + lpk[f]]
            u, inv = torch.unique(v, return_inverse=True)
            uniq_vals.append(u)
            inverse.append(inv + base)
            l = torch.zeros(B, dtype=features.lengths().dtype, device=values.device)
            if B > 0:
                l[0] = u.numel()
            new_lengths.append(l)
            off += lpk[f]
            base += u.numel()
        kjt = KeyedJaggedTensor(keys=features.keys(), values=torch.cat(uniq_vals) if uniq_vals else values, lengths=torch.cat(new_lengths), stride=B)
        return kjt, torch.cat(inverse) if inverse else values

    def compute(self, ctx: EmbeddingCollectionContext, dist_input: KJTList) -> List[torch.Tensor]:
        return [self._engine.lookup(dist_input[0])], dist_input[0]  # type: ignore[return-value]

    def output_dist(self, ctx: EmbeddingCollectionContext, output) -> LazyAwaitable[Dict[str, JaggedTensor]]:
        embs, dist_kjt = output
        emb = embs[0]
        eng = self._engine
        if eng._seq_a2a is not None:
            aw = eng._seq_a2a(emb, dist_kjt.lengths(), ctx.input_splits, ctx.output_splits, batch_size_per_rank=dist_kjt._stride_per_rank)
        else:
            aw = NoWait(emb)

        def finish() -> Dict[str, JaggedTensor]:
            rows = aw.wait()  # rows in routed order: [unit][local sample]
            return self._assemble(ctx, rows)

        return EmbeddingCollectionAwaitable(finish)

    def compute_and_output_dist(self, ctx: EmbeddingCollectionContext, input: KJTList) -> LazyAwaitable[Dict[str, JaggedTensor]]:
        return self.output_dist(ctx, self.compute(ctx, input))

    def _assemble(self, ctx: EmbeddingCollectionContext, rows: torch.Tensor) -> Dict[str, JaggedTensor]:
        """Rows in routed (global unit, sample) order -> per-feature JaggedTensors in the original id order."""
        eng = self._engine
        feats = ctx.features
        assert feats is not None and ctx.routed is not None
        F = len(self._feature_names)
        B = feats.stride()
        lookup_lpk = ctx.lookup_lpk  # ids looked up per feature (== original ids unless deduplicated)
        per_feature: List[torch.Tensor] = []
        if ctx.unbucketize is None:
            # pure key permutation: every unit holds all ids of its feature, column slices side by side
            unit_rows = torch.split(rows, ctx.routed.length_per_key(), dim=0)
            for fi in range(F):
                us = sorted(self._units_of_feature[fi], key=lambda g: eng.units[g].shard.col_off)
                # (rows are as wide as the widest unit of the rank set: a column slice mixed with whole tables is padded on the right)
                parts = [unit_rows[g][:, : eng.units[g].shard.cols] for g in us]
                per_feature.append(parts[0] if len(parts) == 1 else torch.cat(parts, dim=1))
        else:
            # ids were re-sorted by (unit, sample) and replicated once per column slice: undo the sort, then
            # the nc consecutive rows of an id are its column slices in ascending column order
            restored = rows.index_select(0, ctx.unbucketize)
            off = 0
            for fi in range(F):
                slices = sorted({(eng.units[g].shard.col_off, eng.units[g].shard.cols) for g in self._units_of_feature[fi]})
                nc = len(slices)
                n = lookup_lpk[fi]
                block = restored[off : off + n * nc]
                off += n * nc
                if nc > 1:
                    b3 = block.view(n, nc, block.shape[1])
                    block = torch.cat([b3[:, j, :cols] for j, (_, cols) in enumerate(slices)], dim=1)
                else:
                    block = block[:, : slices[0][1]]
                per_feature.append(block)
        lengths2d = feats.lengths().view(F, B)
        values_split = torch.split(feats.values(), feats.length_per_key()) if self._need_indices else None
        out: Dict[str, JaggedTensor] = {}
        rev = ctx.reverse_indices
        pos = base = 0
        orig_lpk = feats.length_per_key()
        for fi, name in enumerate(self._embedding_names):
            emb_f = per_feature[fi]
            if rev is not None:
                emb_f = emb_f.index_select(0, rev[pos : pos + orig_lpk[fi]] - base)
                pos += orig_lpk[fi]
                base += lookup_lpk[fi]
            out[name] = JaggedTensor(values=emb_f, lengths=lengths2d[fi], weights=values_split[fi] if values_split is not None else None)
        return out

    # ---- parameters / state --------------------------------------------------------------------------------------------
    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
        for name, p in self._table_params.items():
            yield (prefix + "." if prefix else "") + f"embeddings.{name}.weight", p

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        yield from ()

    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        for name in self._table_params:
            yield (prefix + "." if prefix else "") + f"embeddings.{name}.weight"
        # the engine's storage (autograd anchor parameters) is rank-private: keep it out of DDP
        for n, _ in nn.Module.named_parameters(self):
            if n.startswith("_engine."):
                yield (prefix + "." if prefix else "") + n

    def state_dict(self, destination: Optional[Dict[str, Any]] = None, prefix: str = "", keep_vars: bool = False) -> Dict[str, Any]:
        from .comm import get_local_size

        if destination is None:
            destination = OrderedDict()
        local_size = get_local_size(self._env.world_size)
        by_table = self._local_shards_by_table()
        for cfg in self._embedding_configs:
            key = f"{prefix}embeddings.{cfg.name}.weight"
            shards = by_table.get(cfg.name, [])
            dtype = torch.float32 if cfg.data_type == DataType.FP32 else data_type_to_dtype(cfg.data_type)
            local = [(w, [s.row_off, s.col_off], [s.rows, s.cols]) for s, w, _, _ in shards]
            if self._pg is None or not dist.is_initialized():
                full = torch.zeros(cfg.num_embeddings, cfg.embedding_dim, dtype=dtype, device=self._device)
                for t, o, s in local:
                    full[o[0] : o[0] + s[0], o[1] : o[1] + s[1]] = t
                destination[key] = full
            else:
                destination[key] = _sharded_view(local, self._global_shards(cfg.name), [cfg.num_embeddings, cfg.embedding_dim], dtype,
                                                 self._plan[cfg.name].sharding_type, self._env, self._pg, self._device.type, local_size)
        return destination

    @torch.no_grad()
    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True, assign: bool = False):
        from torch.distributed._shard.sharded_tensor import ShardedTensor

        missing, unexpected = [], []
        by_table = self._local_shards_by_table()
        expected = set()
        for cfg in self._embedding_configs:
            key = f"embeddings.{cfg.name}.weight"
            expected.add(key)
            if key not in state_dict:
                missing.append(key)
                continue
            src = state_dict[key]
            pieces = _local_pieces_with_offsets(src)
            for s, w, _, _ in by_table.get(cfg.name, []):
                if pieces is not None:
                    for t, off in pieces:
                        if off == [s.row_off, s.col_off]:
                            w.copy_(t)
                else:
                    w.copy_(src[s.row_off : s.row_off + s.rows, s.col_off : s.col_off + s.cols])
        unexpected = [k for k in state_dict.keys() if k not in expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing}, unexpected {unexpected}")
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    @property
    def fused_optimizer(self) -> KeyedOptimizer:
        if self._optim is None:
            self._optim = EmbeddingFusedOptimizer(self, prefix="embeddings")  # type: ignore[arg-type]
        return self._optim

    def embedding_configs(self) -> List[EmbeddingConfig]:
        return self._embedding_configs

    @property
    def unsharded_module_type(self) -> Type[EmbeddingCollection]:
        return EmbeddingCollection


class EmbeddingCollectionSharder(BaseEmbeddingSharder[EmbeddingCollection]):
    def __init__(self, fused_params: Optional[Dict[str, Any]] = None, qcomm_codecs_registry: Optional[Dict[str, QuantizedCommCodecs]] = None,
                 use_index_dedup: bool = False) -> None:
        super().__init__(fused_params, qcomm_codecs_registry)
        self._use_index_dedup = use_index_dedup

    def shard(self, module: EmbeddingCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, device: Optional[torch.device] = None,
              module_fqn: Optional[str] = None) -> ShardedEmbeddingCollection:
        return ShardedEmbeddingCollection(module, params, env, self.fused_params, device, qcomm_codecs_registry=self.qcomm_codecs_registry,
                                          use_index_dedup=self._use_index_dedup, module_fqn=module_fqn)

    def shardable_parameters(self, module: EmbeddingCollection) -> Dict[str, nn.Parameter]:
        return {name.split(".")[0]: param for name, param in module.embeddings.named_parameters()}

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [ShardingType.TABLE_WISE.value, ShardingType.COLUMN_WISE.value, ShardingType.ROW_WISE.value]

    @property
    def module_type(self) -> Type[EmbeddingCollection]:
        return EmbeddingCollection
