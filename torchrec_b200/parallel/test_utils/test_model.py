"""Small reference models for tests (reference ``torchrec/distributed/test_utils/test_model.py``: ``TestDenseArch`` :1088, ``TestOverArch`` :1246,
``TestEBCSparseArch`` :1604, ``TestECSparseArch`` :1562, ``TestSparseNN`` :1748, preproc modules :2123-2470)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from ...modules.embedding_configs import EmbeddingBagConfig, EmbeddingConfig
from ...modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from ...sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor
from .model_input import ModelInput


class TestDenseArch(nn.Module):
    __test__ = False

    def __init__(self, num_float_features: int = 10, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.linear = nn.Linear(num_float_features, 8, device=device)

    def forward(self, dense_input: torch.Tensor) -> torch.Tensor:
        return self.linear(dense_input)


class TestOverArch(nn.Module):
    __test__ = False

    def __init__(self, tables: List[EmbeddingBagConfig], weighted_tables: List[EmbeddingBagConfig], embedding_names: Optional[List[str]] = None,
                 device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._features = [f for t in tables for f in t.feature_names]
        self._weighted_features = [f for t in weighted_tables for f in t.feature_names]
        in_dim = 8 + sum(t.embedding_dim * len(t.feature_names) for t in tables) + sum(t.embedding_dim * len(t.feature_names) for t in weighted_tables)
        self.dhn_arch = nn.Sequential(nn.Linear(in_dim, 16, device=device), nn.ReLU(), nn.Linear(16, 1, device=device))

    def forward(self, dense: torch.Tensor, sparse: KeyedTensor, weighted_sparse: Optional[KeyedTensor] = None) -> torch.Tensor:
        cols = [dense] + [sparse[f] for f in self._features]
        if weighted_sparse is not None:
            cols += [weighted_sparse[f] for f in self._weighted_features]
        return self.dhn_arch(torch.cat(cols, dim=1))


class TestEBCSparseArch(nn.Module):
    __test__ = False

    def __init__(self, tables: List[EmbeddingBagConfig], weighted_tables: List[EmbeddingBagConfig], device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.ebc = EmbeddingBagCollection(tables=tables, device=device)
        self.weighted_ebc = EmbeddingBagCollection(tables=weighted_tables, is_weighted=True, device=device) if weighted_tables else None

    def forward(self, features: KeyedJaggedTensor, weighted_features: Optional[KeyedJaggedTensor] = None) -> Tuple[KeyedTensor, Optional[KeyedTensor]]:
        return self.ebc(features), (self.weighted_ebc(weighted_features) if self.weighted_ebc is not None and weighted_features is not None else None)


class TestECSparseArch(nn.Module):
    __test__ = False

    def __init__(self, tables: List[EmbeddingConfig], device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.ec = EmbeddingCollection(tables=tables, device=device)

    def forward(self, features: KeyedJaggedTensor) -> Dict[str, Any]:
        return self.ec(features)


class TestSparseNNBase(nn.Module):
    __test__ = False

    def __init__(self, tables: List[EmbeddingBagConfig], weighted_tables: Optional[List[EmbeddingBagConfig]] = None, num_float_features: int = 10,
                 dense_device: Optional[torch.device] = None, sparse_device: Optional[torch.device] = None) -> None:
        super().__init__()
        self.dense_device, self.sparse_device = dense_device, sparse_device


class TestSparseNN(TestSparseNNBase):
    """dense MLP + EBC (+ weighted EBC) + over arch; ``forward(ModelInput)`` -> ``(loss, prediction)`` in training, prediction in eval."""

    def __init__(self, tables: List[EmbeddingBagConfig], weighted_tables: Optional[List[EmbeddingBagConfig]] = None, num_float_features: int = 10,
                 dense_device: Optional[torch.device] = None, sparse_device: Optional[torch.device] = None, **_: Any) -> None:
        super().__init__(tables, weighted_tables, num_float_features, dense_device, sparse_device)
        weighted_tables = weighted_tables or []
        self.dense = TestDenseArch(num_float_features, dense_device)
        self.sparse = TestEBCSparseArch(tables, weighted_tables, sparse_device)
        self.over = TestOverArch(tables, weighted_tables, device=dense_device)

    def forward(self, input: ModelInput) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        dense_r = self.dense(input.float_features)
        sparse_r, weighted_r = self.sparse(input.idlist_features, input.idscore_features)
        pred = torch.sigmoid(torch.mean(self.over(dense_r, sparse_r, weighted_r), dim=1))
        if self.training:
            return torch.nn.functional.binary_cross_entropy(pred, input.label), pred
        return pred


class TestPreprocNonWeighted(nn.Module):
    """A pipelineable postproc: clamps ids into range (no parameters)."""

    __test__ = False

    def forward(self, kjt: KeyedJaggedTensor) -> KeyedJaggedTensor:
        return KeyedJaggedTensor(keys=kjt.keys(), values=kjt.values().clamp(min=0), lengths=kjt.lengths(), weights=kjt.weights_or_none(), stride=kjt.stride())


class TestPreprocWeighted(TestPreprocNonWeighted):
    __test__ = False


class TestModelWithPreproc(nn.Module):
    """Sparse inputs pass through parameter-free modules before the EBCs - the case ``pipeline_postproc`` exists for."""

    __test__ = False

    def __init__(self, tables: List[EmbeddingBagConfig], weighted_tables: List[EmbeddingBagConfig], device: torch.device, postproc_module: Optional[nn.Module] = None,
                 num_float_features: int = 10, **_: Any) -> None:
        super().__init__()
        self.dense = TestDenseArch(num_float_features, device)
        self.ebc = EmbeddingBagCollection(tables=tables, device=device)
        self.weighted_ebc = EmbeddingBagCollection(tables=weighted_tables, is_weighted=True, device=device) if weighted_tables else None
        self.postproc_nonweighted = TestPreprocNonWeighted()
        self.postproc_weighted = TestPreprocWeighted()
        self._postproc_module = postproc_module
        self.over = TestOverArch(tables, weighted_tables or [], device=device)

    def forward(self, input: ModelInput) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._postproc_module is not None:
            input = self._postproc_module(input)
        sparse = self.ebc(self.postproc_nonweighted(input.idlist_features))
        weighted = self.weighted_ebc(self.postproc_weighted(input.idscore_features)) if self.weighted_ebc is not None and input.idscore_features is not None else None
        pred = torch.sigmoid(torch.mean(self.over(self.dense(input.float_features), sparse, weighted), dim=1))
        return torch.nn.functional.binary_cross_entropy(pred, input.label), pred


# ---- tower models (reference test_model.py:1888-2120) --------------------------------------------------------------------------------------------------------
class TestTowerInteraction(nn.Module):
    """The interaction of a test tower: a Linear over the concatenated embeddings of the tower's features (unweighted first)."""

    __test__ = False

    def __init__(self, tables: List[EmbeddingBagConfig], weighted_tables: Optional[List[EmbeddingBagConfig]] = None, device: Optional[torch.device] = None) -> None:
        super().__init__()
        self._features = [f for t in tables for f in t.feature_names]
        self._weighted_features = [f for t in (weighted_tables or []) for f in t.feature_names]
        n = sum(t.embedding_dim * len(t.feature_names) for t in list(tables) + list(weighted_tables or []))
        self.linear = nn.Linear(n, n, device=device or torch.device("cpu"))

    def forward(self, sparse: KeyedTensor, weighted_sparse: Optional[KeyedTensor] = None) -> torch.Tensor:
        cols = [sparse[f] for f in self._features]
        if weighted_sparse is not None:
            cols += [weighted_sparse[f] for f in self._weighted_features]
        return self.linear(torch.cat(cols, dim=1))


def _bce_or_pred(model: nn.Module, over_r: torch.Tensor, label: torch.Tensor):
    pred = torch.sigmoid(torch.mean(over_r, dim=1))
    if model.training:
        return torch.nn.functional.binary_cross_entropy_with_logits(pred, label), pred
    return pred


class TestTowerSparseNN(TestSparseNNBase):
    """Dense arch + two ``EmbeddingTower``s (tables 2, 3 and table 0) + a plain sparse arch (table 1 and the first weighted table); needs
    four tables and one weighted table. Towers are sharded as units: all tables of a tower live on one host."""

    def __init__(self, tables: List[EmbeddingBagConfig], num_float_features: int = 10, weighted_tables: Optional[List[EmbeddingBagConfig]] = None,
                 embedding_groups: Optional[Dict[str, List[str]]] = None, dense_device: Optional[torch.device] = None, sparse_device: Optional[torch.device] = None,
                 feature_processor_modules: Optional[Dict[str, nn.Module]] = None) -> None:
        super().__init__(tables, weighted_tables, num_float_features, dense_device, sparse_device)
        from ...modules.embedding_tower import EmbeddingTower

        weighted_tables = weighted_tables or []
        assert len(tables) >= 4 and len(weighted_tables) >= 1, "TestTowerSparseNN needs four tables and one weighted table"
        self.dense = TestDenseArch(num_float_features, dense_device)
        self.tower_0 = EmbeddingTower(EmbeddingBagCollection(tables=[tables[2], tables[3]], device=sparse_device), TestTowerInteraction([tables[2], tables[3]], device=dense_device))
        self.tower_1 = EmbeddingTower(EmbeddingBagCollection(tables=[tables[0]], device=sparse_device), TestTowerInteraction([tables[0]], device=dense_device))
        self.sparse_arch = TestEBCSparseArch([tables[1]], [weighted_tables[0]], sparse_device)
        self._arch_features = list(tables[1].feature_names)
        self._arch_weighted_features = list(weighted_tables[0].feature_names)
        width = 8 + self.tower_0.interaction.linear.out_features + self.tower_1.interaction.linear.out_features \
            + tables[1].embedding_dim * len(tables[1].feature_names) + weighted_tables[0].embedding_dim * len(weighted_tables[0].feature_names)
        self.over = nn.Linear(width, 16, device=dense_device)

    def forward(self, input: ModelInput):
        dense_r = self.dense(input.float_features)
        sparse, weighted = self.sparse_arch(input.idlist_features, input.idscore_features)
        arch = torch.cat([sparse[f] for f in self._arch_features] + [weighted[f] for f in self._arch_weighted_features], dim=1)
        sparse_r = torch.cat([self.tower_0(input.idlist_features), self.tower_1(input.idlist_features), arch], dim=1)
        return _bce_or_pred(self, self.over(torch.cat([dense_r, sparse_r], dim=1)), input.label)


class TestTowerCollectionSparseNN(TestSparseNNBase):
    """Dense arch + one ``EmbeddingTowerCollection`` of three towers: tables 0, 2 / table 1 / the first weighted table (a weighted tower)."""

    def __init__(self, tables: List[EmbeddingBagConfig], num_float_features: int = 10, weighted_tables: Optional[List[EmbeddingBagConfig]] = None,
                 embedding_groups: Optional[Dict[str, List[str]]] = None, dense_device: Optional[torch.device] = None, sparse_device: Optional[torch.device] = None,
                 feature_processor_modules: Optional[Dict[str, nn.Module]] = None) -> None:
        super().__init__(tables, weighted_tables, num_float_features, dense_device, sparse_device)
        from ...modules.embedding_tower import EmbeddingTower, EmbeddingTowerCollection

        weighted_tables = weighted_tables or []
        assert len(tables) >= 3 and len(weighted_tables) >= 1, "TestTowerCollectionSparseNN needs three tables and one weighted table"
        self.dense = TestDenseArch(num_float_features, dense_device)
        t0 = EmbeddingTower(EmbeddingBagCollection(tables=[tables[0], tables[2]], device=sparse_device), TestTowerInteraction([tables[0], tables[2]], device=dense_device))
        t1 = EmbeddingTower(EmbeddingBagCollection(tables=[tables[1]], device=sparse_device), TestTowerInteraction([tables[1]], device=dense_device))
        t2 = EmbeddingTower(EmbeddingBagCollection(tables=[weighted_tables[0]], is_weighted=True, device=sparse_device),
                            TestTowerInteraction([weighted_tables[0]], device=dense_device))
        self.tower_arch = EmbeddingTowerCollection(towers=[t0, t1, t2])
        width = 8 + sum(t.interaction.linear.out_features for t in (t0, t1, t2))
        self.over = nn.Linear(width, 16, device=dense_device)

    def forward(self, input: ModelInput):
        dense_r = self.dense(input.float_features)
        sparse_r = self.tower_arch(input.idlist_features, input.idscore_features)
        return _bce_or_pred(self, self.over(torch.cat([dense_r, sparse_r], dim=1)), input.label)


# ---- pooled + sequence tables in one model (reference test_model.py:2632-2960) -------------------------------------------------------------------------------------
MAX_SEQUENCE_LENGTH = 20
DENSE_LAYER_OUT_SIZE = 8
OVER_ARCH_OUT_SIZE = 16


def _tables_dim_sum(tables: List[Any], per_feature_multiplier: int = 1) -> int:
    return sum(t.embedding_dim * len(t.feature_names) * per_feature_multiplier for t in tables)


class TestMixedSequenceOverArch(nn.Module):
    """One Linear over [dense | pooled embeddings | sequence embeddings padded to ``max_sequence_length``]."""

    __test__ = False

    def __init__(self, ebc_tables: List[EmbeddingBagConfig], ec_tables: List[EmbeddingConfig], weighted_tables: List[EmbeddingBagConfig], device: Optional[torch.device] = None,
                 max_sequence_length: Optional[int] = None, dense_arch_out_size: Optional[int] = None, over_arch_out_size: Optional[int] = None) -> None:
        super().__init__()
        width = (dense_arch_out_size or DENSE_LAYER_OUT_SIZE) + _tables_dim_sum(ebc_tables) + _tables_dim_sum(ec_tables, max_sequence_length or MAX_SEQUENCE_LENGTH) \
            + _tables_dim_sum(weighted_tables)
        self.linear = nn.Linear(width, over_arch_out_size or OVER_ARCH_OUT_SIZE, device=device or torch.device("cpu"))

    def forward(self, dense: torch.Tensor, sparse: torch.Tensor) -> torch.Tensor:
        return self.linear(torch.cat([dense, sparse], dim=1))


class TestMixedSequenceOverArchLargeActivation(nn.Module):
    """The same inputs through a stack of wide hidden layers: a large activation footprint in backward (memory-stashing tests)."""

    __test__ = False

    def __init__(self, ebc_tables: List[EmbeddingBagConfig], ec_tables: List[EmbeddingConfig], weighted_tables: List[EmbeddingBagConfig], device: Optional[torch.device] = None,
                 max_sequence_length: Optional[int] = None, dense_arch_out_size: Optional[int] = None, over_arch_out_size: Optional[int] = None,
                 large_activation_dim: int = 1024, num_hidden_layers: int = 3) -> None:
        super().__init__()
        dev = device or torch.device("cpu")
        width = (dense_arch_out_size or DENSE_LAYER_OUT_SIZE) + _tables_dim_sum(ebc_tables) + _tables_dim_sum(ec_tables, max_sequence_length or MAX_SEQUENCE_LENGTH) \
            + _tables_dim_sum(weighted_tables)
        layers: List[nn.Module] = []
        for i in range(num_hidden_layers):
            layers += [nn.Linear(width if i == 0 else large_activation_dim, large_activation_dim, device=dev), nn.ReLU()]
        layers.append(nn.Linear(large_activation_dim if num_hidden_layers else width, over_arch_out_size or OVER_ARCH_OUT_SIZE, device=dev))
        self.mlp = nn.Sequential(*layers)

    def forward(self, dense: torch.Tensor, sparse: torch.Tensor) -> torch.Tensor:
        return self.mlp(torch.cat([dense, sparse], dim=1))


MIXED_OVER_ARCH_CLASSES: Dict[str, Any] = {"default": TestMixedSequenceOverArch, "large_activation": TestMixedSequenceOverArchLargeActivation}


class TestMixedEmbeddingSparseArch(TestSparseNNBase):
    """Pooled tables (EmbeddingBagConfig -> one EmbeddingBagCollection) and sequence tables (EmbeddingConfig -> one EmbeddingCollection,
    outputs padded to 20 positions) in one model: both collections are sharded and pipelined side by side. ``sparse_forward`` /
    ``dense_forward`` are separate so pipelines can run them in different stages."""

    def __init__(self, tables: List[Any], num_float_features: int = 10, weighted_tables: Optional[List[EmbeddingBagConfig]] = None,
                 embedding_groups: Optional[Dict[str, List[str]]] = None, dense_device: Optional[torch.device] = None, sparse_device: Optional[torch.device] = None,
                 feature_processor_modules: Optional[Dict[str, nn.Module]] = None, over_arch_clazz: Any = TestMixedSequenceOverArch, device: Optional[torch.device] = None,
                 enable_activation_stashing: bool = False, dense_arch_hidden_sizes: Optional[List[int]] = None, over_arch_kwargs: Optional[Dict[str, Any]] = None) -> None:
        super().__init__([], weighted_tables, num_float_features, dense_device, sparse_device)
        dev = device or sparse_device or torch.device("cpu")
        ebc_tables = [t for t in tables if isinstance(t, EmbeddingBagConfig)]
        ec_tables = [t for t in tables if isinstance(t, EmbeddingConfig) and not isinstance(t, EmbeddingBagConfig)]
        if len(ebc_tables) + len(ec_tables) != len(tables):
            raise ValueError(f"Unsupported table type among {[type(t).__name__ for t in tables]}")
        self.ebc = EmbeddingBagCollection(tables=ebc_tables, device=dev) if ebc_tables else None
        self.ec = EmbeddingCollection(tables=ec_tables, device=dev) if ec_tables else None
        self.ec_embedding_dim = ec_tables[0].embedding_dim if ec_tables else 0
        self._ebc_features = [f for t in ebc_tables for f in t.feature_names]
        self._ec_features = [f for t in ec_tables for f in t.feature_names]
        self._enable_activation_stashing = enable_activation_stashing
        hidden = list(dense_arch_hidden_sizes or [])
        if hidden:
            layers: List[nn.Module] = []
            for i, h in enumerate(hidden):
                layers += [nn.Linear(num_float_features if i == 0 else hidden[i - 1], h, device=dense_device), nn.ReLU()]
            self.dense: nn.Module = nn.Sequential(*layers)
            dense_out = hidden[-1]
        else:
            self.dense = TestDenseArch(num_float_features, dense_device)
            dense_out = DENSE_LAYER_OUT_SIZE
        if isinstance(over_arch_clazz, str):
            over_arch_clazz = MIXED_OVER_ARCH_CLASSES[over_arch_clazz]
        self.over = over_arch_clazz(ebc_tables, ec_tables, weighted_tables or [], dense_device, dense_arch_out_size=dense_out, **(over_arch_kwargs or {}))

    def sparse_forward(self, input: ModelInput) -> torch.Tensor:
        from ...ops import jagged as J

        features = input.idlist_features
        parts: List[torch.Tensor] = []
        if self.ebc is not None:
            kt = self.ebc(features if features.keys() == self._ebc_features else features.permute([features.keys().index(f) for f in self._ebc_features]))
            parts.append(kt.values())
        if self.ec is not None:
            out = self.ec(features if features.keys() == self._ec_features else features.permute([features.keys().index(f) for f in self._ec_features]))
            for f in self._ec_features:
                jt = out[f]
                parts.append(J.jagged_2d_to_dense(jt.values(), jt.offsets(), MAX_SEQUENCE_LENGTH).reshape(-1, MAX_SEQUENCE_LENGTH * self.ec_embedding_dim))
        return torch.cat(parts, dim=1)

    def dense_forward(self, input: ModelInput, sparse_output: torch.Tensor):
        return _bce_or_pred(self, self.over(self.dense(input.float_features), sparse_output), input.label)

    def forward(self, input: ModelInput):
        return self.dense_forward(input, self.sparse_forward(input))
