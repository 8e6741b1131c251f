"""Sharded quantized (inference) embedding collections — single process driving several GPUs
(reference torchrec/distributed/quant_embeddingbag.py:171, quant_embedding.py:597, sharding/tw_sharding.py:466-584).

Inference has no process group (``ShardingEnv.from_local``): the module owns one quantized kernel per device.
Input dist = split the KJT per device + P2P copies (``KJTOneToAll``); output dist = NVLink copies back to the
first device, concatenated (TW/CW) or summed (RW)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Type

import torch
from torch import nn

from ..modules.embedding_configs import EmbeddingBagConfig, EmbeddingConfig, PoolingType
from ..modules.embedding_modules import get_embedding_names_by_table
from ..ops.quant_tbe import QuantTableBatchedEmbeddingBags
from ..quant.embedding_modules import EmbeddingBagCollection as QuantEmbeddingBagCollection
from ..quant.embedding_modules import EmbeddingCollection as QuantEmbeddingCollection
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from .embedding_types import BaseQuantEmbeddingSharder
from .engine import shards_of
from .types import NullShardedModuleContext, ParameterSharding, ShardedModule, ShardingEnv, ShardingType, LazyNoWait, NoWait


def _dev(device_type: str, rank: int) -> torch.device:
    return torch.device("cpu") if device_type != "cuda" else torch.device("cuda", rank)


class ShardedQuantEmbeddingBagCollection(ShardedModule[List[KeyedJaggedTensor], List[torch.Tensor], KeyedTensor, NullShardedModuleContext]):
    """TW / CW / RW sharded quantized EBC for inference."""

    def __init__(self, module: QuantEmbeddingBagCollection, table_name_to_parameter_sharding: Dict[str, ParameterSharding], env: ShardingEnv,
                 fused_params: Optional[Dict[str, Any]] = None, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._env = env
        W = env.world_size
        self._device_type = "cuda" if (device is not None and device.type == "cuda") or (device is None and torch.cuda.is_available()) else "cpu"
        tables = module.embedding_bag_configs()
        self._tables = tables
        self._is_weighted = module.is_weighted()
        self._output_dtype = module.output_dtype()
        self._embedding_names = [n for ns in get_embedding_names_by_table(tables) for n in ns]
        self._dims = [t.embedding_dim for t in tables for _ in t.feature_names]
        self._feature_names = [f for t in tables for f in t.feature_names]
        feat_table = [ti for ti, t in enumerate(tables) for _ in t.feature_names]
        out_base = [0]
        for d in self._dims:
            out_base.append(out_base[-1] + d)
        self._total_cols = out_base[-1]
        src_weights = {t.name: module.embedding_bags[t.name].weight for t in tables}
        # per rank: list of (table idx, shard, feature idx)
        per_rank: List[List[Tuple[int, Any, int]]] = [[] for _ in range(W)]
        self._row_sharded_feature: Dict[int, List[Tuple[int, int, int]]] = {}
        for ti, t in enumerate(tables):
            ps = table_name_to_parameter_sharding[t.name]
            for s in shards_of(ti, t, ps):
                for fi in [i for i, x in enumerate(feat_table) if x == ti]:
                    per_rank[s.rank].append((ti, s, fi))
        self._per_rank = per_rank
        self._tbes = nn.ModuleList()
        self._rank_meta: List[List[Tuple[int, Any, int]]] = []
        for r in range(W):
            units = per_rank[r]
            dev = _dev(self._device_type, r)
            if not units:
                self._tbes.append(nn.Identity())
                self._rank_meta.append([])
                continue
            shard_keys: List[Tuple[int, int, int]] = []
            specs = []
            for ti, s, fi in units:
                key = (ti, s.row_off, s.col_off)
                if key not in shard_keys:
                    shard_keys.append(key)
                    specs.append((f"{tables[ti].name}_{s.row_off}_{s.col_off}", s.rows, s.cols, tables[ti].data_type))
            fmap = [shard_keys.index((ti, s.row_off, s.col_off)) for ti, s, fi in units]
            pool = {tables[ti].pooling for ti, _, _ in units}
            tbe = QuantTableBatchedEmbeddingBags(specs, fmap, pooling_mode=0, output_dtype=self._output_dtype, device=dev)
            # fill: dequantise the source slice and re-quantise the shard (column slices change the row scale)
            from ..ops.quant_tbe import dequantize_rows

            done = set()
            for (ti, s, fi), k in zip(units, fmap):
                if k in done:
                    continue
                done.add(k)
                full = dequantize_rows(src_weights[tables[ti].name].to("cpu"), tables[ti].embedding_dim, tables[ti].data_type)
                tbe.assign_from_float(k, full[s.row_off : s.row_off + s.rows, s.col_off : s.col_off + s.cols])
            self._tbes.append(tbe)
            self._rank_meta.append(units)
        self._mean_features = [fi for fi, ti in enumerate(feat_table) if tables[ti].pooling == PoolingType.MEAN]
        self._out_base = out_base

    def create_context(self) -> NullShardedModuleContext:
        return NullShardedModuleContext()

    def input_dist(self, ctx, features: KeyedJaggedTensor):
        """Per device: the features of its units with ids filtered / rebased to the shard's row range."""
        pos = {k: i for i, k in enumerate(features.keys())}
        B = features.stride()
        out: List[Optional[KeyedJaggedTensor]] = []
        jt = features.to_dict()
        for r, units in enumerate(self._rank_meta):
            if not units:
                out.append(None)
                continue
            dev = _dev(self._device_type, r)
            vals, lens, ws = [], [], []
            for ti, s, fi in units:
                f = jt[self._feature_names[fi]]
                v, l = f.values(), f.lengths()
                w = f.weights_or_none()
                if s.rows != self._tables[ti].num_embeddings:
                    keep = (v >= s.row_off) & (v < s.row_off + s.rows)
                    seg = torch.repeat_interleave(torch.arange(B, device=v.device), l.long())
                    l = torch.zeros(B, dtype=l.dtype, device=v.device).index_add_(0, seg[keep], torch.ones(int(keep.sum()), dtype=l.dtype, device=v.device))
                    v = v[keep] - s.row_off
                    w = w[keep] if w is not None else None
                vals.append(v)
                lens.append(l)
                if w is not None:
                    ws.append(w)
            kjt = KeyedJaggedTensor(keys=[self._feature_names[fi] for _, _, fi in units], values=torch.cat(vals), lengths=torch.cat(lens),
                                    weights=torch.cat(ws) if ws else None, stride=B)
            out.append(kjt.to(dev, non_blocking=True))
        self._divisor = None
        if self._mean_features:
            lengths = features.lengths().view(len(features.keys()), B)
            self._divisor = {fi: lengths[pos[self._feature_names[fi]]].clamp(min=1).float() for fi in self._mean_features}
        return NoWait(NoWait(out))

    def compute(self, ctx, dist_input) -> List[Optional[torch.Tensor]]:
        outs = []
        for r, kjt in enumerate(dist_input):
            if kjt is None:
                outs.append(None)
                continue
            psw = kjt.weights_or_none() if self._is_weighted else None
            outs.append(self._tbes[r](kjt.values(), kjt.offsets(), psw, batch_size=kjt.stride()))
        return outs

    def output_dist(self, ctx, output):
        dev0 = _dev(self._device_type, 0)
        B = next(o.shape[0] for o in output if o is not None)
        res = torch.zeros(B, self._total_cols, dtype=self._output_dtype, device=dev0)
        for r, o in enumerate(output):
            if o is None:
                continue
            o = o.to(dev0, non_blocking=True)
            c = 0
            for ti, s, fi in self._rank_meta[r]:
                base = self._out_base[fi] + s.col_off
                res[:, base : base + s.cols] += o[:, c : c + s.cols]
                c += s.cols
        if self._divisor:
            for fi, d in self._divisor.items():
                res[:, self._out_base[fi] : self._out_base[fi + 1]] /= d.to(dev0).unsqueeze(1)
        return LazyNoWait(KeyedTensor(keys=self._embedding_names, length_per_key=self._dims, values=res))

    def forward(self, features: KeyedJaggedTensor):
        ctx = self.create_context()
        return self.output_dist(ctx, self.compute(ctx, self.input_dist(ctx, features).wait().wait())).wait()

    @property
    def unsharded_module_type(self):
        return QuantEmbeddingBagCollection


class QuantEmbeddingBagCollectionSharder(BaseQuantEmbeddingSharder[QuantEmbeddingBagCollection]):
    def shard(self, module, params, env, device=None, module_fqn=None) -> ShardedQuantEmbeddingBagCollection:
        return ShardedQuantEmbeddingBagCollection(module, params, env, self.fused_params, device=device)

    def shardable_parameters(self, module: QuantEmbeddingBagCollection) -> Dict[str, nn.Parameter]:
        return {name: nn.Parameter(torch.empty(t.num_embeddings, t.embedding_dim, device="meta"), requires_grad=False)
                for name, t in ((t.name, t) for t in module.embedding_bag_configs())}

    @property
    def module_type(self) -> Type[QuantEmbeddingBagCollection]:
        return QuantEmbeddingBagCollection
