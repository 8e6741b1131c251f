"""Model selection config (reference ``torchrec/distributed/test_utils/model_config.py``: ``BaseModelConfig`` :45, per-model configs :92-298,
``create_model_config`` :301, ``ModelSelectionConfig`` :327)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, field, fields
from typing import Any, Dict, List, Optional

import torch
from torch import nn

from ...modules.embedding_modules import EmbeddingBagCollection


@dataclass
class BaseModelConfig(ABC):
    num_float_features: int = 10

    @abstractmethod
    def generate_model(self, tables: List[Any], weighted_tables: List[Any], dense_device: torch.device, **kwargs: Any) -> nn.Module:
        ...


@dataclass
class TestSparseNNConfig(BaseModelConfig):
    __test__ = False

    def generate_model(self, tables, weighted_tables, dense_device, **kwargs):
        from .test_model import TestSparseNN

        return TestSparseNN(tables=tables, weighted_tables=weighted_tables, num_float_features=self.num_float_features, dense_device=dense_device, sparse_device=torch.device("meta"))


@dataclass
class TestModelWithPreprocConfig(BaseModelConfig):
    __test__ = False

    def generate_model(self, tables, weighted_tables, dense_device, **kwargs):
        from .test_model import TestModelWithPreproc

        return TestModelWithPreproc(tables, weighted_tables, dense_device, num_float_features=self.num_float_features)


@dataclass
class DLRMConfig(BaseModelConfig):
    embedding_dim: int = 128
    dense_arch_layer_sizes: List[int] = field(default_factory=lambda: [512, 256, 128])
    over_arch_layer_sizes: List[int] = field(default_factory=lambda: [1024, 1024, 512, 256, 1])

    def generate_model(self, tables, weighted_tables, dense_device, **kwargs):
        from ...models.dlrm import DLRM, DLRMTrain

        ebc = EmbeddingBagCollection(tables=tables, device=torch.device("meta"))
        return _BatchAdapter(DLRMTrain(DLRM(ebc, self.num_float_features, self.dense_arch_layer_sizes, self.over_arch_layer_sizes, dense_device=dense_device)))


@dataclass
class TestTowerSparseNNConfig(BaseModelConfig):
    """Two embedding towers + a plain sparse arch (``TestTowerSparseNN``: four tables, one weighted table)."""

    __test__ = False
    embedding_groups: Optional[Dict[str, List[str]]] = None
    feature_processor_modules: Optional[Dict[str, nn.Module]] = None

    def generate_model(self, tables, weighted_tables, dense_device, **kwargs: Any) -> nn.Module:
        from .test_model import TestTowerSparseNN

        return TestTowerSparseNN(tables=tables, num_float_features=self.num_float_features, weighted_tables=weighted_tables, dense_device=dense_device,
                                 sparse_device=torch.device("meta"), embedding_groups=self.embedding_groups, feature_processor_modules=self.feature_processor_modules)


@dataclass
class TestTowerCollectionSparseNNConfig(BaseModelConfig):
    """One tower collection of three towers (``TestTowerCollectionSparseNN``: three tables, one weighted table)."""

    __test__ = False
    embedding_groups: Optional[Dict[str, List[str]]] = None
    feature_processor_modules: Optional[Dict[str, nn.Module]] = None

    def generate_model(self, tables, weighted_tables, dense_device, **kwargs: Any) -> nn.Module:
        from .test_model import TestTowerCollectionSparseNN

        return TestTowerCollectionSparseNN(tables=tables, num_float_features=self.num_float_features, weighted_tables=weighted_tables, dense_device=dense_device,
                                           sparse_device=torch.device("meta"), embedding_groups=self.embedding_groups, feature_processor_modules=self.feature_processor_modules)


@dataclass
class DeepFMConfig(BaseModelConfig):
    hidden_layer_size: int = 20
    deep_fm_dimension: int = 5

    def generate_model(self, tables, weighted_tables, dense_device, **kwargs):
        from ...models.deepfm import SimpleDeepFMNN

        ebc = EmbeddingBagCollection(tables=tables, device=torch.device("meta"))
        return _BatchAdapter(_WithLoss(SimpleDeepFMNN(self.num_float_features, ebc, self.hidden_layer_size, self.deep_fm_dimension)))


@dataclass
class MixedEmbeddingConfig(BaseModelConfig):
    """Pooled + sequence tables in one model (``TestMixedEmbeddingSparseArch``); ``over_arch_clazz``: a class or ``"default"`` /
    ``"large_activation"``."""

    embedding_groups: Optional[Dict[str, List[str]]] = None
    over_arch_clazz: Any = "default"
    enable_activation_stashing: bool = False
    dense_arch_hidden_sizes: Optional[List[int]] = None
    over_arch_kwargs: Optional[Dict[str, Any]] = None

    def __post_init__(self) -> None:
        from .test_model import MIXED_OVER_ARCH_CLASSES

        if isinstance(self.over_arch_clazz, str):
            if self.over_arch_clazz not in MIXED_OVER_ARCH_CLASSES:
                raise ValueError(f"Unknown mixed over_arch_clazz: {self.over_arch_clazz}. Available: {list(MIXED_OVER_ARCH_CLASSES.keys())}")
            self.over_arch_clazz = MIXED_OVER_ARCH_CLASSES[self.over_arch_clazz]

    def generate_model(self, tables, weighted_tables=None, dense_device=None, **kwargs: Any) -> nn.Module:
        from .test_model import TestMixedEmbeddingSparseArch

        return TestMixedEmbeddingSparseArch(tables=tables, num_float_features=self.num_float_features, weighted_tables=weighted_tables, embedding_groups=self.embedding_groups,
                                            dense_device=dense_device, sparse_device=torch.device("meta"), over_arch_clazz=self.over_arch_clazz, device=torch.device("meta"),
                                            enable_activation_stashing=self.enable_activation_stashing, dense_arch_hidden_sizes=self.dense_arch_hidden_sizes,
                                            over_arch_kwargs=self.over_arch_kwargs)


class _WithLoss(nn.Module):
    def __init__(self, model: nn.Module) -> None:
        super().__init__()
        self.model = model

    def forward(self, batch: Any):
        logits = self.model(batch.dense_features, batch.sparse_features).squeeze(-1)
        return torch.nn.functional.binary_cross_entropy_with_logits(logits, batch.labels.float()), (logits.detach(), batch.labels)


class _BatchAdapter(nn.Module):
    """``ModelInput`` -> the (dense_features, sparse_features, labels) batch the model families expect."""

    def __init__(self, model: nn.Module) -> None:
        super().__init__()
        self.model = model

    def forward(self, input: Any):
        from ...datasets.utils import Batch

        return self.model(Batch(dense_features=input.float_features, sparse_features=input.idlist_features, labels=input.label))


_REGISTRY = {"test_sparse_nn": TestSparseNNConfig, "test_model_with_preproc": TestModelWithPreprocConfig, "dlrm": DLRMConfig, "deepfm": DeepFMConfig,
             "test_tower_sparse_nn": TestTowerSparseNNConfig, "test_tower_collection_sparse_nn": TestTowerCollectionSparseNNConfig, "mixed_embedding": MixedEmbeddingConfig}


def create_model_config(model_name: str, **kwargs: Any) -> BaseModelConfig:
    if model_name not in _REGISTRY:
        raise ValueError(f"unknown model {model_name!r}; available: {sorted(_REGISTRY)}")
    cls = _REGISTRY[model_name]
    names = {f.name for f in fields(cls)}
    return cls(**{k: v for k, v in kwargs.items() if k in names})


@dataclass
class ModelSelectionConfig:
    model_name: str = "test_sparse_nn"
    num_float_features: int = 10
    embedding_dim: int = 128
    extra: Dict[str, Any] = field(default_factory=dict)

    def get_model_config(self) -> BaseModelConfig:
        return create_model_config(self.model_name, num_float_features=self.num_float_features, embedding_dim=self.embedding_dim, **self.extra)

    def create_test_model(self, tables: List[Any], weighted_tables: List[Any], dense_device: torch.device) -> nn.Module:
        return self.get_model_config().generate_model(tables, weighted_tables, dense_device)
