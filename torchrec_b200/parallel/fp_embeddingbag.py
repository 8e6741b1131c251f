"""Sharded FeatureProcessedEmbeddingBagCollection (reference torchrec/distributed/fp_embeddingbag.py:64-294).

The feature processors run AFTER the input dist, on the rank that hosts the table (so a position-weight vector is only
touched where its feature lives); their parameters stay data-parallel (DDP-wrapped by DistributedModelParallel). The
per-sample-weight gradient comes from ``csrc/tbe_psw_grad.cu``. For row-sharded tables the position inside the bag is
lost by the id bucketing, so the processors are applied before the input dist on the source rank instead."""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Type, Union

import torch
from torch import nn

from ..modules.feature_processor_ import FeatureProcessorsCollection
from ..modules.fp_embedding_modules import FeatureProcessedEmbeddingBagCollection
from ..sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from .embedding_types import BaseEmbeddingSharder, KJTList
from .embeddingbag import EmbeddingBagCollectionContext, EmbeddingBagCollectionSharder, ShardedEmbeddingBagCollection
from .types import Awaitable, LazyAwaitable, ParameterSharding, ShardedModule, ShardingEnv, ShardingType


def _position_table(processors, is_collection: bool) -> Optional[Dict[str, nn.Parameter]]:
    """feature -> position-weight vector when every processor is position weighted, else None."""
    from ..modules.feature_processor_ import PositionWeightedModule, PositionWeightedModuleCollection

    if is_collection:
        inner = processors
        if hasattr(inner, "_feature_processors"):  # FeatureProcessorDictWrapper
            inner = inner._feature_processors
            if all(isinstance(m, PositionWeightedModule) for m in inner.values()):
                return {k: m.position_weight for k, m in inner.items()}
            return None
        return dict(inner.position_weights.items()) if isinstance(inner, PositionWeightedModuleCollection) else None
    if all(isinstance(m, PositionWeightedModule) for m in processors.values()):
        return {k: m.position_weight for k, m in processors.items()}
    return None


def positions_into_weights(features: KeyedJaggedTensor) -> KeyedJaggedTensor:
    """Store the position of every id inside its bag in the weights slot, so it survives id bucketing and the input
    dist (reference fp_embeddingbag.py:117-122 / feature_processor_.py modify_input_for_feature_processor)."""
    from ..ops import jagged as J

    pos = J.offsets_range(features.offsets()[:-1].long(), features.values().numel()).float()
    return KeyedJaggedTensor(keys=features.keys(), values=features.values(), weights=pos, lengths=features.lengths(), offsets=features.offsets(),
                             stride=features.stride(), length_per_key=features.length_per_key())


def weights_from_positions(features: KeyedJaggedTensor, table: Dict[str, nn.Parameter]) -> KeyedJaggedTensor:
    """weights[i] = position_weight[key(i)][position(i)] for a KJT whose weights slot carries positions (keys may repeat)."""
    keys, lpk = features.keys(), features.length_per_key()
    pos = torch.split(features.weights().long(), lpk)
    out = []
    for k, p in zip(keys, pos):
        if k in table:
            out.append(torch.gather(table[k], 0, p.clamp(max=table[k].numel() - 1)))
        else:
            out.append(torch.ones(p.numel(), device=p.device))
    return KeyedJaggedTensor(keys=keys, values=features.values(), lengths=features.lengths(), offsets=features.offsets(), weights=torch.cat(out) if out else features.weights(),
                             stride=features.stride(), stride_per_rank=features._stride_per_rank, length_per_key=lpk)


def apply_feature_processors_by_position(features: KeyedJaggedTensor, processors: Union[nn.ModuleDict, FeatureProcessorsCollection], is_collection: bool) -> KeyedJaggedTensor:
    """Apply per-feature processors to a KJT whose keys may repeat (one key per lookup unit after the input dist)."""
    keys = features.keys()
    if len(keys) == 0:  # this rank hosts no table of the group
        return features
    if len(set(keys)) == len(keys) and is_collection and features._stride_per_rank is None:
        return processors(features)
    lpk = features.length_per_key()
    vals = torch.split(features.values(), lpk)
    lens = features.lengths().view(len(keys), -1)
    w = features.weights_or_none()
    ws = torch.split(w, lpk) if w is not None else [None] * len(keys)
    out_w: List[torch.Tensor] = []
    for i, k in enumerate(keys):
        if is_collection:
            out_w.append(processors(KeyedJaggedTensor(keys=[k], values=vals[i], lengths=lens[i], weights=ws[i], stride=lens[i].numel())).weights())
        elif k in processors:
            out_w.append(processors[k](JaggedTensor(values=vals[i], lengths=lens[i], weights=ws[i])).weights())
        else:
            out_w.append(ws[i] if ws[i] is not None else torch.ones(vals[i].numel(), device=vals[i].device))
    return KeyedJaggedTensor(keys=keys, values=features.values(), lengths=features.lengths(), offsets=features.offsets(), weights=torch.cat(out_w), stride=features.stride(),
                             stride_per_rank=features._stride_per_rank, length_per_key=lpk)


def param_dp_sync(kt: KeyedTensor, no_op_tensor: torch.Tensor) -> KeyedTensor:
    """Tie a (zero) tensor computed from every data-parallel processor parameter into the output, so that all of them take part in the
    backward on every rank (DDP expects a gradient for each of its parameters even when a rank saw none of a feature's ids)."""
    kt._values.add_(no_op_tensor)
    return kt


class ShardedFeatureProcessedEmbeddingBagCollection(ShardedModule):
    def __init__(self, module: FeatureProcessedEmbeddingBagCollection, table_name_to_parameter_sharding: Dict[str, ParameterSharding],
                 ebc_sharder: EmbeddingBagCollectionSharder, env: ShardingEnv, device: torch.device, module_fqn: Optional[str] = None) -> None:
        super().__init__()
        self._device, self._env = device, env
        self._embedding_bag_collection: ShardedEmbeddingBagCollection = ebc_sharder.shard(module._embedding_bag_collection, table_name_to_parameter_sharding, env=env, device=device)
        self._embedding_bag_collection._needs_dist_kjt = True  # processors run on the distributed KJT (jagged view + weight gradients)
        self._row_wise_sharded = any(ps.sharding_type in (ShardingType.ROW_WISE.value, ShardingType.TABLE_ROW_WISE.value, ShardingType.GRID_SHARD.value)
                                     for ps in table_name_to_parameter_sharding.values())
        self._has_dp = any(ps.sharding_type == ShardingType.DATA_PARALLEL.value for ps in table_name_to_parameter_sharding.values())
        fp = module._feature_processors
        self._is_collection = isinstance(fp, FeatureProcessorsCollection)
        self._feature_processors = fp if self._is_collection else nn.ModuleDict(dict(fp.items()))
        self._feature_processors.to(device)
        # row-sharded tables bucket ids by row block, which destroys the in-bag position: ship the positions instead
        self._pos_table = _position_table(self._feature_processors, self._is_collection)
        self._positions_mode = self._row_wise_sharded and self._pos_table is not None
        if self._row_wise_sharded and self._pos_table is None:
            raise NotImplementedError("row-wise sharded feature-processed tables need position-weighted processors "
                                      "(any other processor would see bucketed bags); use table/column-wise sharding")
        self.register_buffer("_no_op_zero", torch.zeros(1, device=device), persistent=False)

    def create_context(self) -> EmbeddingBagCollectionContext:
        return self._embedding_bag_collection.create_context()

    def input_dist(self, ctx, features: KeyedJaggedTensor):
        if self._positions_mode:
            features = positions_into_weights(features)
        return self._embedding_bag_collection.input_dist(ctx, features)

    def _one(self, f: KeyedJaggedTensor) -> KeyedJaggedTensor:
        if self._positions_mode:
            return weights_from_positions(f, self._pos_table)
        return apply_feature_processors_by_position(f, self._feature_processors, self._is_collection)

    def _fp(self, ctx, dist_input: KJTList) -> KJTList:
        if getattr(ctx, "dp_features", None) is not None:  # replicated tables look up the local batch
            ctx.dp_features = self._one(ctx.dp_features)
        return KJTList([self._one(f) for f in dist_input])

    def compute(self, ctx, dist_input: KJTList):
        return self._embedding_bag_collection.compute(ctx, self._fp(ctx, dist_input))

    def output_dist(self, ctx, output) -> LazyAwaitable[KeyedTensor]:
        return self._sync(self._embedding_bag_collection.output_dist(ctx, output))

    def compute_and_output_dist(self, ctx, input: KJTList) -> LazyAwaitable[KeyedTensor]:
        return self._sync(self._embedding_bag_collection.compute_and_output_dist(ctx, self._fp(ctx, input)))

    def _sync(self, aw: LazyAwaitable[KeyedTensor]) -> LazyAwaitable[KeyedTensor]:
        """Every processor parameter joins the autograd graph on every rank, so DDP's bucket order never diverges
        (a rank that hosts no table of a feature would otherwise never produce that gradient)."""
        params = [p.flatten() for p in self._feature_processors.parameters()]
        if not params:
            return aw
        no_op = self._no_op_zero * torch.cat(params).sum()

        def cb(kt: KeyedTensor) -> KeyedTensor:
            return KeyedTensor(keys=kt.keys(), length_per_key=kt.length_per_key(), values=kt.values() + no_op.to(kt.values().dtype), key_dim=kt.key_dim())

        aw.callbacks.append(cb)
        return aw

    def sharded_parameter_names(self, prefix: str = "") -> Iterator[str]:
        p = prefix + "." if prefix else ""
        yield from self._embedding_bag_collection.sharded_parameter_names(p + "_embedding_bag_collection")

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        from .types import delegating_named_parameters

        yield from delegating_named_parameters(self, prefix, recurse)

    @property
    def fused_optimizer(self):
        return self._embedding_bag_collection.fused_optimizer


class FeatureProcessedEmbeddingBagCollectionSharder(BaseEmbeddingSharder[FeatureProcessedEmbeddingBagCollection]):
    def __init__(self, ebc_sharder: Optional[EmbeddingBagCollectionSharder] = None, fused_params=None, qcomm_codecs_registry=None) -> None:
        super().__init__(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)
        self._ebc_sharder = ebc_sharder or EmbeddingBagCollectionSharder(fused_params=fused_params, qcomm_codecs_registry=qcomm_codecs_registry)

    def shard(self, module: FeatureProcessedEmbeddingBagCollection, params: Dict[str, ParameterSharding], env: ShardingEnv,
              device: Optional[torch.device] = None, module_fqn: Optional[str] = None) -> ShardedFeatureProcessedEmbeddingBagCollection:
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return ShardedFeatureProcessedEmbeddingBagCollection(module, params, self._ebc_sharder, env, device, module_fqn)

    @property
    def fused_params(self):
        return self._ebc_sharder.fused_params

    def shardable_parameters(self, module: FeatureProcessedEmbeddingBagCollection) -> Dict[str, nn.Parameter]:
        return self._ebc_sharder.shardable_parameters(module._embedding_bag_collection)

    @property
    def module_type(self) -> Type[FeatureProcessedEmbeddingBagCollection]:
        return FeatureProcessedEmbeddingBagCollection
