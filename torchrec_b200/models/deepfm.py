"""DeepFM model (reference torchrec/models/deepfm.py:20-400): sparse arch (EBC) + dense projection + FM / deep-FM
interaction + over arch."""
from typing import List

import torch
from torch import nn

from ..modules.deepfm import DeepFM, FactorizationMachine
from ..modules.embedding_modules import EmbeddingBagCollection
from ..sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor


class SparseArch(nn.Module):
    def __init__(self, embedding_bag_collection: EmbeddingBagCollection) -> None:
        super().__init__()
        self.embedding_bag_collection = embedding_bag_collection

    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        return self.embedding_bag_collection(features)


class DenseArch(nn.Module):
    """dense features [B, in] -> [B, embedding_dim] (same space as the sparse embeddings)."""

    def __init__(self, in_features: int, hidden_layer_size: int, embedding_dim: int) -> None:
        super().__init__()
        self.model = nn.Sequential(nn.Linear(in_features, hidden_layer_size), nn.ReLU(), nn.Linear(hidden_layer_size, embedding_dim), nn.ReLU())

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        return self.model(features)


class FMInteractionArch(nn.Module):
    """cat(dense, deep_fm(dense + sparse), fm(dense + sparse)) -> [B, D + deep_fm_dimension + 1]."""

    def __init__(self, fm_in_features: int, sparse_feature_names: List[str], deep_fm_dimension: int) -> None:
        super().__init__()
        self.sparse_feature_names = sparse_feature_names
        self.deep_fm = DeepFM(dense_module=nn.Sequential(nn.Linear(fm_in_features, deep_fm_dimension), nn.ReLU()))
        self.fm = FactorizationMachine()

    def forward(self, dense_features: torch.Tensor, sparse_features: KeyedTensor) -> torch.Tensor:
        if len(self.sparse_feature_names) == 0:
            return dense_features
        tensor_list: List[torch.Tensor] = [dense_features]
        d = sparse_features.to_dict()
        for feature_name in self.sparse_feature_names:
            tensor_list.append(d[feature_name].to(dense_features.dtype))
        return torch.cat([dense_features, self.deep_fm(tensor_list), self.fm(tensor_list)], dim=1)


class OverArch(nn.Module):
    def __init__(self, in_features: int) -> None:
        super().__init__()
        self.model = nn.Sequential(nn.Linear(in_features, 1), nn.Sigmoid())

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        return self.model(features)


class SimpleDeepFMNN(nn.Module):
    def __init__(self, num_dense_features: int, embedding_bag_collection: EmbeddingBagCollection, hidden_layer_size: int, deep_fm_dimension: int) -> None:
        super().__init__()
        cfgs = embedding_bag_collection.embedding_bag_configs()
        assert len(cfgs) > 0, "At least one embedding bag is required"
        for i in range(len(cfgs)):
            assert cfgs[i].embedding_dim == cfgs[0].embedding_dim, "All EmbeddingBagConfigs must have the same dimension"
        embedding_dim: int = cfgs[0].embedding_dim
        feature_names = []
        fm_in_features = embedding_dim
        for conf in cfgs:
            for feat in conf.feature_names:
                feature_names.append(feat)
                fm_in_features += conf.embedding_dim
        self.sparse_arch = SparseArch(embedding_bag_collection)
        self.dense_arch = DenseArch(in_features=num_dense_features, hidden_layer_size=hidden_layer_size, embedding_dim=embedding_dim)
        self.inter_arch = FMInteractionArch(fm_in_features=fm_in_features, sparse_feature_names=feature_names, deep_fm_dimension=deep_fm_dimension)
        self.over_arch = OverArch(embedding_dim + deep_fm_dimension + 1)

    def forward(self, dense_features: torch.Tensor, sparse_features: KeyedJaggedTensor) -> torch.Tensor:
        embedded_dense = self.dense_arch(dense_features)
        embedded_sparse = self.sparse_arch(sparse_features)
        return self.over_arch(self.inter_arch(dense_features=embedded_dense, sparse_features=embedded_sparse))


class SimpleDeepFMNNWrapper(SimpleDeepFMNN):
    """Inference-friendly twin whose forward returns a ``{task: prediction}`` dict."""

    def forward(self, dense_features: torch.Tensor, sparse_features: KeyedJaggedTensor):  # type: ignore[override]
        return {"default": super().forward(dense_features, sparse_features).squeeze(-1)}
