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
