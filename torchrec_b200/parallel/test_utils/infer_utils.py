"""Helpers for inference tests: build a small model, quantize it, shard the quantized model over local devices and compare
(reference ``torchrec/distributed/test_utils/infer_utils.py``: ``TestModelInfo`` :96, ``quantize`` :321, ``create_test_model`` :630,
``shard_qebc`` :853, ``shard_qec`` :926, ``assert_close`` :979, ``assert_weight_spec`` :1005, mock TBEs :1042).

Everything builds on this framework's own pieces: ``quantize_embeddings`` (module swap to the quantized collections),
``shard_quant_model`` (inference planner + sharding over the devices of one process) and the quantized TBE kernels."""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Type, Union

import torch
from torch import nn

from ...inference.modules import quantize_embeddings, set_pruning_data, shard_quant_model
from ...modules.embedding_configs import DataType, EmbeddingBagConfig, EmbeddingConfig
from ...modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from ...modules.fp_embedding_modules import FeatureProcessedEmbeddingBagCollection
from ...quant.embedding_modules import EmbeddingBagCollection as QuantEmbeddingBagCollection
from ...quant.embedding_modules import EmbeddingCollection as QuantEmbeddingCollection
from ...quant.embedding_modules import quant_prep_enable_quant_state_dict_split_scale_bias, quant_prep_enable_register_tbes
from ...sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from ..planner import EmbeddingShardingPlanner, Topology
from ..planner.enumerators import EmbeddingEnumerator, EmbeddingPerfEstimator, EmbeddingStorageEstimator
from ..planner.types import ParameterConstraints
from ..quant_embedding import QuantEmbeddingCollectionSharder
from ..quant_embeddingbag import QuantEmbeddingBagCollectionSharder
from ..types import ShardingPlan, ShardingType
from .model_input import ModelInput
from .test_model import TestSparseNN

_WEIGHT_DTYPES = {torch.qint8: DataType.INT8, torch.quint8: DataType.INT8, torch.int8: DataType.INT8, torch.uint8: DataType.INT8, torch.quint4x2: DataType.INT4,
                  torch.float16: DataType.FP16, torch.float32: DataType.FP32, torch.bfloat16: DataType.BF16}


def _data_type(dtype: Union[torch.dtype, DataType]) -> DataType:
    return dtype if isinstance(dtype, DataType) else _WEIGHT_DTYPES[dtype]


@dataclass
class TestModelInfo:
    __test__ = False

    sparse_device: torch.device
    dense_device: torch.device
    num_features: int
    num_float_features: int
    num_weighted_features: int
    tables: Union[List[EmbeddingBagConfig], List[EmbeddingConfig]] = field(default_factory=list)
    weighted_tables: List[EmbeddingBagConfig] = field(default_factory=list)
    model: nn.Module = field(default_factory=nn.Module)
    quant_model: nn.Module = field(default_factory=nn.Module)
    sharders: List[Any] = field(default_factory=list)
    topology: Optional[Topology] = None
    planner: Optional[EmbeddingShardingPlanner] = None


# ---- input plumbing: modules that take plain tensors instead of a KJT / ModelInput (tracing, export, serving fronts) -----------------------------
class KJTInputWrapper(nn.Module):
    """``forward(keys, values, weights, lengths, offsets)`` -> ``module(KeyedJaggedTensor(...))``."""

    def __init__(self, module_kjt_input: nn.Module) -> None:
        super().__init__()
        self._module_kjt_input = module_kjt_input

    def forward(self, keys: List[str], values: torch.Tensor, weights: Optional[torch.Tensor] = None, lengths: Optional[torch.Tensor] = None,
                offsets: Optional[torch.Tensor] = None):
        return self._module_kjt_input(KeyedJaggedTensor(keys=keys, values=values, weights=weights, lengths=lengths, offsets=offsets))


class KJTInputExportWrapper(nn.Module):
    """The keys are fixed at construction (export traces tensors only): ``forward(values, lengths, weights=None)``; a KeyedTensor /
    dict-of-JaggedTensor result is flattened to a list of tensors."""

    def __init__(self, module_kjt_input: nn.Module, kjt_keys: List[str]) -> None:
        super().__init__()
        self._module_kjt_input = module_kjt_input
        self._kjt_keys = list(kjt_keys)

    def _kjt(self, values, lengths, weights=None, **kw) -> KeyedJaggedTensor:
        return KeyedJaggedTensor(keys=self._kjt_keys, values=values, lengths=lengths, weights=weights, **kw)

    @staticmethod
    def _flatten(out):
        if isinstance(out, KeyedTensor):
            return [out.values()]
        if isinstance(out, dict):
            return [t for jt in out.values() for t in ((jt.values(), jt.lengths()) if isinstance(jt, JaggedTensor) else (jt,))]
        return out

    def forward(self, values: torch.Tensor, lengths: torch.Tensor, weights: Optional[torch.Tensor] = None):
        return self._flatten(self._module_kjt_input(self._kjt(values, lengths, weights)))


class KJTInputExportDynamicShapeWrapper(KJTInputExportWrapper):
    """Same, for exports with a dynamic number of values (the output is returned unflattened)."""

    def forward(self, values: torch.Tensor, lengths: torch.Tensor, weights: Optional[torch.Tensor] = None):
        return self._module_kjt_input(self._kjt(values, lengths, weights))


class KJTInputExportWrapperWithStrides(KJTInputExportWrapper):
    """Variable batch per feature: ``stride_per_key_per_rank`` travels as a tensor argument."""

    def forward(self, values: torch.Tensor, lengths: torch.Tensor, stride_per_key_per_rank: torch.Tensor, weights: Optional[torch.Tensor] = None):
        return self._flatten(self._module_kjt_input(self._kjt(values, lengths, weights, stride_per_key_per_rank=stride_per_key_per_rank.tolist())))


class TorchTypesModelInputWrapper(nn.Module):
    """A model that takes a ``ModelInput`` behind a signature of plain tensors / key lists (scripting and serving fronts cannot pass the
    dataclass); ``model_input_to_forward_args`` produces the arguments."""

    def __init__(self, module: nn.Module) -> None:
        super().__init__()
        self._module = module

    def copy(self, device: torch.device) -> "TorchTypesModelInputWrapper":
        return copy.deepcopy(self).to(device)

    def forward(self, float_features: torch.Tensor, idlist_features_keys: List[str], idlist_features_values: torch.Tensor, idscore_features_keys: List[str],
                idscore_features_values: torch.Tensor, idscore_features_weights: torch.Tensor, label: torch.Tensor, idlist_features_lengths: Optional[torch.Tensor] = None,
                idlist_features_offsets: Optional[torch.Tensor] = None, idscore_features_lengths: Optional[torch.Tensor] = None,
                idscore_features_offsets: Optional[torch.Tensor] = None):
        idlist = KeyedJaggedTensor(keys=idlist_features_keys, values=idlist_features_values, lengths=idlist_features_lengths, offsets=idlist_features_offsets)
        idscore = None
        if idscore_features_keys:
            idscore = KeyedJaggedTensor(keys=idscore_features_keys, values=idscore_features_values, weights=idscore_features_weights, lengths=idscore_features_lengths,
                                        offsets=idscore_features_offsets)
        return self._module(ModelInput(float_features=float_features, idlist_features=idlist, idscore_features=idscore, label=label))


def model_input_to_forward_args_kjt(mi: ModelInput) -> Tuple[List[str], torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Arguments of ``KJTInputWrapper.forward`` for the unweighted features of a batch."""
    kjt = mi.idlist_features
    assert isinstance(kjt, KeyedJaggedTensor)
    return kjt.keys(), kjt.values(), kjt.weights_or_none(), kjt.lengths_or_none(), kjt.offsets_or_none()


def model_input_to_forward_args(mi: ModelInput) -> Tuple[Any, ...]:
    """Arguments of ``TorchTypesModelInputWrapper.forward`` (positional order of its signature)."""
    idlist, idscore = mi.idlist_features, mi.idscore_features
    assert isinstance(idlist, KeyedJaggedTensor)
    empty_i, empty_f = torch.zeros(0, dtype=torch.int64, device=idlist.values().device), torch.zeros(0, device=idlist.values().device)
    return (mi.float_features, idlist.keys(), idlist.values(), idscore.keys() if idscore is not None else [], idscore.values() if idscore is not None else empty_i,
            idscore.weights() if idscore is not None else empty_f, mi.label, idlist.lengths_or_none(), idlist.offsets_or_none(),
            idscore.lengths_or_none() if idscore is not None else None, idscore.offsets_or_none() if idscore is not None else None)


def prep_inputs(model_info: TestModelInfo, world_size: int, batch_size: int = 1, count: int = 5, long_indices: bool = True) -> List[ModelInput]:
    """``count`` random batches for the tables of ``model_info`` (ids int64 or int32)."""
    return [ModelInput.generate(batch_size=batch_size, tables=model_info.tables, weighted_tables=model_info.weighted_tables, num_float_features=model_info.num_float_features,
                                indices_dtype=torch.int64 if long_indices else torch.int32, lengths_dtype=torch.int64 if long_indices else torch.int32).to(model_info.sparse_device)
            for _ in range(count)]


def prep_inputs_multiprocess(model_info: TestModelInfo, world_size: int, batch_size: int = 1, count: int = 5) -> List[Tuple[ModelInput, ...]]:
    """Per step one batch per rank."""
    return [tuple(ModelInput.generate_local_batches(world_size, batch_size, tables=model_info.tables, weighted_tables=model_info.weighted_tables,
                                                    num_float_features=model_info.num_float_features)) for _ in range(count)]


# ---- quantize ---------------------------------------------------------------------------------------------------------------------------------------------------------
def create_cw_min_partition_constraints(table_min_partition_pairs: List[Tuple[str, int]]) -> Dict[str, ParameterConstraints]:
    return {name: ParameterConstraints(sharding_types=[ShardingType.COLUMN_WISE.value], min_partition=mp) for name, mp in table_min_partition_pairs}


def quantize(module: nn.Module, inplace: bool, output_type: torch.dtype = torch.float, register_tbes: bool = False, quant_state_dict_split_scale_bias: bool = False,
             weight_dtype: Union[torch.dtype, DataType] = torch.qint8, per_table_weight_dtypes: Optional[Dict[str, Union[torch.dtype, DataType]]] = None) -> nn.Module:
    """Swap every EmbeddingBagCollection / EmbeddingCollection of ``module`` for its quantized version (row-wise ``weight_dtype``, per-table
    overrides); ``register_tbes`` exposes the kernels as ``tbes``; ``quant_state_dict_split_scale_bias`` splits scale / bias in the state dict."""
    if not inplace:
        module = copy.deepcopy(module)
    if register_tbes:
        quant_prep_enable_register_tbes(module, [EmbeddingBagCollection, EmbeddingCollection])
    if quant_state_dict_split_scale_bias:
        for m in module.modules():
            if isinstance(m, (EmbeddingBagCollection, EmbeddingCollection)):
                quant_prep_enable_quant_state_dict_split_scale_bias(m)
    per_table = {k: _data_type(v) for k, v in per_table_weight_dtypes.items()} if per_table_weight_dtypes else None
    return quantize_embeddings(module, dtype=_data_type(weight_dtype), inplace=True, output_dtype=output_type, per_table_weight_dtype=per_table)


def quantize_fpebc(module: nn.Module, inplace: bool, output_type: torch.dtype = torch.float, register_tbes: bool = False, quant_state_dict_split_scale_bias: bool = False,
                   weight_dtype: Union[torch.dtype, DataType] = torch.qint8, per_table_weight_dtypes: Optional[Dict[str, Union[torch.dtype, DataType]]] = None) -> nn.Module:
    """``quantize`` for models with feature-processed bags: the embedding bags inside every FeatureProcessedEmbeddingBagCollection are
    quantized, the (float) feature processors stay."""
    if not inplace:
        module = copy.deepcopy(module)
    for m in module.modules():
        if isinstance(m, FeatureProcessedEmbeddingBagCollection):
            m._embedding_bag_collection = quantize(m._embedding_bag_collection, True, output_type, register_tbes, quant_state_dict_split_scale_bias, weight_dtype,
                                                   per_table_weight_dtypes)
    return module


# ---- sharders that force one sharding type / kernel (the planner then has no choice) --------------------------------------------------------------------------
class _FixedChoice:
    def __init__(self, sharding_type: str, kernel_type: str, fused_params: Optional[Dict[str, Any]] = None, shardable_params: Optional[List[str]] = None) -> None:
        super().__init__(fused_params=fused_params, shardable_params=shardable_params)  # type: ignore[call-arg]
        self._sharding_type, self._kernel_type = sharding_type, kernel_type

    def sharding_types(self, compute_device_type: str) -> List[str]:
        return [self._sharding_type]

    def compute_kernels(self, sharding_type: str, compute_device_type: str) -> List[str]:
        return [self._kernel_type]

    def shardable_parameters(self, module: nn.Module) -> Dict[str, nn.Parameter]:
        params = super().shardable_parameters(module)  # type: ignore[misc]
        return {k: v for k, v in params.items() if not self._shardable_params or k in self._shardable_params}  # type: ignore[attr-defined]


class TestQuantEBCSharder(_FixedChoice, QuantEmbeddingBagCollectionSharder):
    __test__ = False


class TestQuantECSharder(_FixedChoice, QuantEmbeddingCollectionSharder):
    __test__ = False


class TestQuantFPEBCSharder(TestQuantEBCSharder):
    """Feature-processed quantized bags shard like plain quantized bags (the processors are replicated)."""

    __test__ = False


# ---- models ------------------------------------------------------------------------------------------------------------------------------------------------------------
def _inference_planner(topology: Topology, batch_size: int, constraints: Optional[Dict[str, ParameterConstraints]]) -> EmbeddingShardingPlanner:
    from ..planner.storage_reservations import FixedPercentageStorageReservation

    enumerator = EmbeddingEnumerator(topology=topology, batch_size=batch_size, constraints=constraints,
                                     estimator=[EmbeddingPerfEstimator(topology=topology, constraints=constraints, is_inference=True),
                                                EmbeddingStorageEstimator(topology=topology, constraints=constraints, is_inference=True)])
    return EmbeddingShardingPlanner(topology=topology, batch_size=batch_size, enumerator=enumerator, constraints=constraints,
                                    storage_reservation=FixedPercentageStorageReservation(percentage=0.0))


def _tables(cls, n: int, rows: int, dim: int, prefix: str, feat_prefix: str):
    return [cls(num_embeddings=rows, embedding_dim=dim, name=f"{prefix}{i}", feature_names=[f"{feat_prefix}{i}"]) for i in range(n)]


def create_test_model(num_embeddings: int, emb_dim: int, world_size: int, batch_size: int, dense_device: torch.device, sparse_device: torch.device,
                      quant_state_dict_split_scale_bias: bool = False, num_features: int = 1, num_float_features: int = 8, num_weighted_features: int = 1,
                      constraints: Optional[Dict[str, ParameterConstraints]] = None, weight_dtype: Union[torch.dtype, DataType] = torch.qint8,
                      pruning_dict: Optional[Dict[str, int]] = None) -> TestModelInfo:
    """TestSparseNN (dense + bags + weighted bags + over arch) behind the plain-tensor signature, in eval mode, with its quantized copy
    and an inference planner over ``world_size`` local devices."""
    topology = Topology(world_size=world_size, local_world_size=world_size, compute_device=sparse_device.type)
    mi = TestModelInfo(dense_device=dense_device, sparse_device=sparse_device, num_features=num_features, num_float_features=num_float_features,
                       num_weighted_features=num_weighted_features, topology=topology, planner=_inference_planner(topology, batch_size, constraints))
    mi.tables = _tables(EmbeddingBagConfig, num_features, num_embeddings, emb_dim, "table_", "feature_")
    mi.weighted_tables = _tables(EmbeddingBagConfig, num_weighted_features, num_embeddings, emb_dim, "weighted_table_", "weighted_feature_")
    if pruning_dict:
        for cfg in mi.tables + mi.weighted_tables:
            if cfg.name in pruning_dict:
                cfg.num_embeddings_post_pruning = pruning_dict[cfg.name]
    mi.model = TorchTypesModelInputWrapper(TestSparseNN(tables=mi.tables, weighted_tables=mi.weighted_tables, num_float_features=num_float_features,
                                                        dense_device=dense_device, sparse_device=sparse_device))
    mi.model.train(False)
    if pruning_dict:
        set_pruning_data(mi.model, pruning_dict)
    mi.quant_model = quantize(mi.model, inplace=False, quant_state_dict_split_scale_bias=quant_state_dict_split_scale_bias, weight_dtype=weight_dtype)
    return mi


class _EBCOnly(nn.Module):
    def __init__(self, ebc: nn.Module) -> None:
        super().__init__()
        self.sparse = nn.Module()
        self.sparse.ebc = ebc

    def forward(self, kjt: KeyedJaggedTensor):
        return self.sparse.ebc(kjt)


def create_test_model_ebc_only_no_quantize(num_embeddings: int, emb_dim: int, world_size: int, batch_size: int, dense_device: torch.device, sparse_device: torch.device,
                                           num_features: int = 1, num_float_features: int = 8, num_weighted_features: int = 1, compute_device: str = "cuda",
                                           feature_processor: bool = False) -> TestModelInfo:
    """Only the embedding bags (optionally behind position weights), not quantized yet: ``model`` is ``KJTInputWrapper(_EBCOnly)`` with the
    collection at ``_module_kjt_input.sparse.ebc``."""
    topology = Topology(world_size=world_size, local_world_size=world_size, compute_device=compute_device if (compute_device != "cuda" or torch.cuda.is_available()) else "cpu")
    mi = TestModelInfo(dense_device=dense_device, sparse_device=sparse_device, num_features=num_features, num_float_features=num_float_features,
                       num_weighted_features=num_weighted_features, topology=topology, planner=_inference_planner(topology, batch_size, None))
    mi.tables = _tables(EmbeddingBagConfig, num_features, num_embeddings, emb_dim, "table_", "feature_")
    mi.weighted_tables = []
    ebc: nn.Module = EmbeddingBagCollection(tables=mi.tables, is_weighted=feature_processor, device=sparse_device)
    if feature_processor:
        from ...modules.feature_processor_ import PositionWeightedModuleCollection

        ebc = FeatureProcessedEmbeddingBagCollection(ebc, PositionWeightedModuleCollection({f: 100 for t in mi.tables for f in t.feature_names}, device=sparse_device))
    mi.model = KJTInputWrapper(_EBCOnly(ebc))
    mi.model.train(False)
    return mi


def create_test_model_ebc_only(num_embeddings: int, emb_dim: int, world_size: int, batch_size: int, dense_device: torch.device, sparse_device: torch.device,
                               num_features: int = 1, num_float_features: int = 8, num_weighted_features: int = 1, quant_state_dict_split_scale_bias: bool = False,
                               compute_device: str = "cuda", feature_processor: bool = False) -> TestModelInfo:
    mi = create_test_model_ebc_only_no_quantize(num_embeddings, emb_dim, world_size, batch_size, dense_device, sparse_device, num_features, num_float_features,
                                                num_weighted_features, compute_device, feature_processor)
    q = quantize_fpebc if feature_processor else quantize
    mi.quant_model = q(mi.model, inplace=False, register_tbes=True, quant_state_dict_split_scale_bias=quant_state_dict_split_scale_bias)
    return mi


# ---- shard ----------------------------------------------------------------------------------------------------------------------------------------------------------------
def _check_expected(plan: ShardingPlan, fqn: str, names: List[str], sharding_type: ShardingType,
                    expected_shards: Optional[List[List[Tuple[Tuple[int, int, int, int], str]]]]) -> None:
    if expected_shards is None:
        return
    msp = plan.plan[fqn]
    for i, name in enumerate(names):
        ps = msp[name]
        assert ps.sharding_type == sharding_type.value, (name, ps.sharding_type)
        assert ps.sharding_spec is not None and len(ps.sharding_spec.shards) == len(expected_shards[i]), (name, ps.sharding_spec)
        for shard, ((off_r, off_c, size_r, size_c), placement) in zip(ps.sharding_spec.shards, expected_shards[i]):
            assert list(shard.shard_offsets) == [off_r, off_c] and list(shard.shard_sizes) == [size_r, size_c], (name, shard)
            assert str(shard.placement) == placement, (name, str(shard.placement), placement)


def _shard(mi: TestModelInfo, sharder: Any, sharding_type: ShardingType, device: torch.device, expected_shards, plan: Optional[ShardingPlan], fqn: str,
           names: List[str]) -> nn.Module:
    if plan is None:
        assert mi.planner is not None
        plan = mi.planner.plan(mi.quant_model, [sharder])
    _check_expected(plan, fqn, names, sharding_type, expected_shards)
    assert mi.topology is not None
    sharded, _ = shard_quant_model(copy.deepcopy(mi.quant_model), world_size=mi.topology.world_size, compute_device=device.type, sharding_device=str(device),
                                   sharders=[sharder], sharding_plan=plan)
    return sharded


def shard_qebc(mi: TestModelInfo, sharding_type: ShardingType, device: torch.device,
               expected_shards: Optional[List[List[Tuple[Tuple[int, int, int, int], str]]]] = None, plan: Optional[ShardingPlan] = None,
               ebc_fqn: str = "_module.sparse.ebc", shard_score_ebc: bool = False, feature_processor: bool = False) -> nn.Module:
    """A sharded copy of ``mi.quant_model`` (left untouched for comparison) with every bag table placed as ``sharding_type``; optionally
    asserts the shards (offsets, sizes, placement) the plan chose."""
    from ..embedding_types import EmbeddingComputeKernel

    names = [t.name for t in mi.tables] + ([t.name for t in mi.weighted_tables] if shard_score_ebc or feature_processor else [])
    cls = TestQuantFPEBCSharder if feature_processor else TestQuantEBCSharder
    sharder = cls(sharding_type=sharding_type.value, kernel_type=EmbeddingComputeKernel.QUANT.value, shardable_params=names)
    return _shard(mi, sharder, sharding_type, device, expected_shards, plan, ebc_fqn, [t.name for t in mi.tables])


def shard_qec(mi: TestModelInfo, sharding_type: ShardingType, device: torch.device,
              expected_shards: Optional[List[List[Tuple[Tuple[int, int, int, int], str]]]] = None, plan: Optional[ShardingPlan] = None,
              ec_fqn: str = "_module_kjt_input.0") -> nn.Module:
    from ..embedding_types import EmbeddingComputeKernel

    sharder = TestQuantECSharder(sharding_type=sharding_type.value, kernel_type=EmbeddingComputeKernel.QUANT.value, shardable_params=[t.name for t in mi.tables])
    return _shard(mi, sharder, sharding_type, device, expected_shards, plan, ec_fqn, [t.name for t in mi.tables])


# ---- comparisons -----------------------------------------------------------------------------------------------------------------------------------------------------------
def assert_close(expected: Any, actual: Any) -> None:
    """Recursive comparison of model outputs: dicts, lists / tuples, JaggedTensor / KeyedJaggedTensor / KeyedTensor, tensors (compared on
    the host, default tolerances)."""
    if isinstance(expected, dict):
        assert list(expected.keys()) == list(actual.keys()), (list(expected.keys()), list(actual.keys()))
        for k in expected:
            assert_close(expected[k], actual[k])
    elif isinstance(expected, (list, tuple)):
        assert len(expected) == len(actual)
        for a, b in zip(expected, actual):
            assert_close(a, b)
    elif isinstance(expected, (JaggedTensor, KeyedJaggedTensor)):
        assert_close(expected.values(), actual.values())
        assert_close(expected.lengths(), actual.lengths())
        assert_close(expected.weights_or_none(), actual.weights_or_none())
        if isinstance(expected, KeyedJaggedTensor):
            assert expected.keys() == actual.keys()
    elif isinstance(expected, KeyedTensor):
        assert expected.keys() == actual.keys() and expected.length_per_key() == actual.length_per_key()
        assert_close(expected.values(), actual.values())
    elif expected is None:
        assert actual is None
    else:
        assert isinstance(expected, torch.Tensor) and isinstance(actual, torch.Tensor), (type(expected), type(actual))
        torch.testing.assert_close(actual.detach().cpu().float() if actual.is_floating_point() else actual.cpu(),
                                   expected.detach().cpu().float() if expected.is_floating_point() else expected.cpu())


def assert_weight_spec(weights_spec: Dict[str, Any], all_expected_shards: List[List[Tuple[Tuple[int, int, int, int], str]]], ebc_fqn: str, weights_prefix: str,
                       all_table_names: List[str], sharding_type: str) -> None:
    """The ``sharded_module_weights_spec`` of a sharded quantized model (fqn of every shard tensor -> table fqn, offsets, sizes, sharding
    type) against the expected shards. Column-wise shards are named ``<table>_<i>``, the others keep the table name."""
    col = sharding_type == ShardingType.COLUMN_WISE.value
    for suffix in ("weight", "weight_qscale", "weight_qbias"):
        for table_name, expected in zip(all_table_names, all_expected_shards):
            for i, ((off_r, off_c, size_r, size_c), placement) in enumerate(expected):
                shard_name = f"{table_name}_{i}" if col else table_name
                key = f"{ebc_fqn}.tbes.{i if not col else 0}.{shard_name}.{suffix}"
                matches = [k for k in weights_spec if k.startswith(ebc_fqn) and k.endswith(f"{shard_name}.{suffix}")]
                if not matches and suffix != "weight":
                    continue  # fused scale / bias: no separate tensors
                assert matches, f"no weight spec entry for {key}; have {list(weights_spec)[:8]}"
                spec = weights_spec[matches[0]]
                fqn, offsets, sizes, st = spec[0], list(spec[1]), list(spec[2]), spec[3]
                assert fqn == f"{ebc_fqn}.{weights_prefix}.{table_name}.{suffix}", (fqn, table_name)
                assert offsets[0] == off_r and sizes[0] == size_r and st == sharding_type, (spec, expected)
                if suffix == "weight":
                    assert offsets[1] == off_c, (spec, expected)


# ---- mock kernels: shape-correct stand-ins for fast plan / spec tests -----------------------------------------------------------------------------------------------
class MockTBE(nn.Module):
    """Stands in for a quantized table-batched kernel: remembers the table specs, returns zeros of the right shape."""

    def __init__(self, embedding_specs: List[Tuple[str, int, int, Any, Any]], device: torch.device, output_dtype: torch.dtype = torch.float32,
                 pooling_mode: Any = None, **kwargs: Any) -> None:
        super().__init__()
        self.embedding_specs = embedding_specs
        self.output_dtype = output_dtype
        self.pooling_mode = pooling_mode
        self.current_device = device
        self.weights_physical_offsets = [0] * len(embedding_specs)
        self.kwargs = kwargs

    def forward(self, indices: torch.Tensor, offsets: torch.Tensor, per_sample_weights: Optional[torch.Tensor] = None, batch_size: Optional[int] = None,
                **_: Any) -> torch.Tensor:
        T = max(len(self.embedding_specs), 1)
        B = batch_size if batch_size is not None else max((offsets.numel() - 1) // T, 0)
        pooled = self.pooling_mode is None or str(self.pooling_mode).upper().rsplit(".", 1)[-1] != "NONE"
        if pooled:
            return torch.zeros(B, sum(s[2] for s in self.embedding_specs), dtype=self.output_dtype, device=indices.device)
        return torch.zeros(indices.numel(), max((s[2] for s in self.embedding_specs), default=0), dtype=self.output_dtype, device=indices.device)

    def split_embedding_weights(self, split_scale_shifts: bool = True) -> List[Tuple[torch.Tensor, Optional[torch.Tensor]]]:
        return [(torch.zeros(rows, dim, dtype=torch.uint8), None) for _, rows, dim, _, _ in self.embedding_specs]


def mock_tbe_from_tbe(tbe: nn.Module) -> MockTBE:
    specs = [(s[0], s[1], s[2], s[3], None) if len(s) == 4 else tuple(s) for s in getattr(tbe, "embedding_specs", [])]  # (name, rows, dim, dtype[, location])
    device = getattr(tbe, "current_device", None) or next((b.device for b in tbe.buffers()), torch.device("cpu"))
    return MockTBE(specs, device, getattr(tbe, "output_dtype", torch.float32), getattr(tbe, "pooling_mode", None))


def _is_quant_kernel(m: nn.Module) -> bool:
    return type(m).__name__ in ("QuantTableBatchedEmbeddingBags", "IntNBitTableBatchedEmbeddingBagsCodegen", "QuantBatchedEmbeddingBag", "QuantBatchedEmbedding")


def replace_registered_tbes_with_mock_tbes(M: nn.Module, path: str = "") -> None:
    """Every quantized kernel registered as a sub-module of ``M`` (``tbes`` lists of ``register_tbes``) becomes a ``MockTBE``."""
    for name, child in list(M.named_children()):
        if _is_quant_kernel(child):
            setattr(M, name, mock_tbe_from_tbe(child))
        else:
            replace_registered_tbes_with_mock_tbes(child, f"{path}.{name}" if path else name)


def replace_sharded_quant_modules_tbes_with_mock_tbes(M: nn.Module) -> None:
    """Same for the kernels the sharded quantized collections hold in plain lists (per device)."""
    for m in M.modules():
        for attr in ("_tbes", "_lookups"):
            lst = getattr(m, attr, None)
            if isinstance(lst, list):
                for i, k in enumerate(lst):
                    if isinstance(k, nn.Module) and _is_quant_kernel(k):
                        lst[i] = mock_tbe_from_tbe(k)
    replace_registered_tbes_with_mock_tbes(M)
