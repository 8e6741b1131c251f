"""Sharded quantized EmbeddingCollection for inference (table-wise placement over local devices;
reference torchrec/distributed/quant_embedding.py:597)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Type

import torch
from torch import nn

from ..quant.embedding_modules import EmbeddingCollection as QuantEmbeddingCollection
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor
from .embedding_types import BaseQuantEmbeddingSharder
from .engine import shards_of
from .types import LazyNoWait, NoWait, NullShardedModuleContext, ParameterSharding, ShardedModule, ShardingEnv, ShardingType


class ShardedQuantEmbeddingCollection(ShardedModule[List[KeyedJaggedTensor], List[torch.Tensor], Dict[str, JaggedTensor], NullShardedModuleContext]):
    def __init__(self, module: QuantEmbeddingCollection, params: Dict[str, ParameterSharding], env: ShardingEnv, fused_params=None, device=None) -> None:
        super().__init__()
        from ..ops.quant_tbe import QuantTableBatchedEmbeddingBags

        self._device_type = "cuda" if torch.cuda.is_available() and (device is None or device.type == "cuda") else "cpu"
        tables = module.embedding_configs()
        self._tables = tables
        self._need_indices = module.need_indices()
        self._out_dtype = module.output_dtype()
        W = env.world_size
        self._rank_tables: List[List[int]] = [[] for _ in range(W)]
        for ti, t in enumerate(tables):
            ps = params[t.name]
            if ps.sharding_type != ShardingType.TABLE_WISE.value:
                raise NotImplementedError("quantized EmbeddingCollection inference sharding supports table_wise placement")
            self._rank_tables[shards_of(ti, t, ps)[0].rank].append(ti)
        self._tbes = nn.ModuleList()
        for r in range(W):
            idxs = self._rank_tables[r]
            dev = torch.device("cuda", r) if self._device_type == "cuda" else torch.device("cpu")
            if not idxs:
                self._tbes.append(nn.Identity())
                continue
            specs = [(tables[i].name, tables[i].num_embeddings, tables[i].embedding_dim, tables[i].data_type) for i in idxs]
            fmap = [k for k, i in enumerate(idxs) for _ in tables[i].feature_names]
            tbe = QuantTableBatchedEmbeddingBags(specs, fmap, pooling_mode=2, output_dtype=self._out_dtype, device=dev)
            for k, i in enumerate(idxs):
                tbe.split_embedding_weights()[k].copy_(module.embeddings[tables[i].name].weight.to(dev))
            self._tbes.append(tbe)

    def create_context(self):
        return NullShardedModuleContext()

    def input_dist(self, ctx, features: KeyedJaggedTensor):
        out = []
        for r, idxs in enumerate(self._rank_tables):
            if not idxs:
                out.append(None)
                continue
            feats = [f for i in idxs for f in self._tables[i].feature_names]
            sub = features.permute([features.keys().index(f) for f in feats])
            dev = torch.device("cuda", r) if self._device_type == "cuda" else torch.device("cpu")
            out.append(sub.to(dev, non_blocking=True))
        return NoWait(NoWait(out))

    def compute(self, ctx, dist_input):
        return [None if k is None else (self._tbes[r](k.values(), k.offsets(), None, batch_size=k.stride()), k) for r, k in enumerate(dist_input)]

    def output_dist(self, ctx, output):
        dev0 = torch.device("cuda", 0) if self._device_type == "cuda" else torch.device("cpu")
        res: Dict[str, JaggedTensor] = {}
        for item in output:
            if item is None:
                continue
            emb, kjt = item
            parts = torch.split(emb.to(dev0), kjt.length_per_key(), dim=0)
            vals = torch.split(kjt.values().to(dev0), kjt.length_per_key()) if self._need_indices else None
            lengths = kjt.lengths().to(dev0).view(len(kjt.keys()), kjt.stride())
            for i, k in enumerate(kjt.keys()):
                res[k] = JaggedTensor(values=parts[i], lengths=lengths[i], weights=vals[i] if vals is not None else None)
        return LazyNoWait(res)

    def forward(self, features: KeyedJaggedTensor):
        ctx = self.create_context()
        return self.output_dist(ctx, self.compute(ctx, self.input_dist(ctx, features).wait().wait())).wait()


class QuantEmbeddingCollectionSharder(BaseQuantEmbeddingSharder[QuantEmbeddingCollection]):
    def shard(self, module, params, env, device=None, module_fqn=None):
        return ShardedQuantEmbeddingCollection(module, params, env, self.fused_params, device)

    def shardable_parameters(self, module: QuantEmbeddingCollection) -> Dict[str, nn.Parameter]:
        return {t.name: nn.Parameter(torch.empty(t.num_embeddings, t.embedding_dim, device="meta"), requires_grad=False) for t in module.embedding_configs()}

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [ShardingType.TABLE_WISE.value]

    @property
    def module_type(self) -> Type[QuantEmbeddingCollection]:
        return QuantEmbeddingCollection
