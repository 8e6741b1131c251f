"""FeatureProcessedEmbeddingBagCollection: weights produced by feature processors feed a weighted EBC
(reference torchrec/modules/fp_embedding_modules.py:68)."""
from typing import Dict, List, Set, Tuple, Union

import torch
from torch import nn

from ..sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor
from .embedding_modules import EmbeddingBagCollection
from .feature_processor_ import FeatureProcessor, FeatureProcessorsCollection


def apply_feature_processors_to_kjt(features: KeyedJaggedTensor, feature_processors: Dict[str, nn.Module]) -> KeyedJaggedTensor:
    processed_weights = []
    features_dict = features.to_dict()
    for key in features.keys():
        jt = features_dict[key]
        if key in feature_processors:
            fp_jt = feature_processors[key](jt)
            processed_weights.append(fp_jt.weights())
        else:
            processed_weights.append(torch.ones(jt.values().shape[0], device=jt.values().device))
    return KeyedJaggedTensor(keys=features.keys(), values=features.values(), weights=torch.cat(processed_weights) if processed_weights else None,
                             lengths=features.lengths(), offsets=features._offsets, stride=features._stride, length_per_key=features._length_per_key,
                             offset_per_key=features._offset_per_key, index_per_key=features._index_per_key)


class FeatureProcessorDictWrapper(FeatureProcessorsCollection):
    def __init__(self, feature_processors: nn.ModuleDict) -> None:
        super().__init__()
        self._feature_processors = feature_processors

    def forward(self, features: KeyedJaggedTensor) -> KeyedJaggedTensor:
        return apply_feature_processors_to_kjt(features, self._feature_processors)


class FeatureProcessedEmbeddingBagCollection(nn.Module):
    """EBC whose per-id weights come from (position-weighted) feature processors."""

    def __init__(self, embedding_bag_collection: EmbeddingBagCollection, feature_processors: Union[Dict[str, FeatureProcessor], FeatureProcessorsCollection]) -> None:
        super().__init__()
        self._embedding_bag_collection = embedding_bag_collection
        self._feature_processors: Union[nn.ModuleDict, FeatureProcessorsCollection]
        if isinstance(feature_processors, FeatureProcessorsCollection):
            self._feature_processors = feature_processors
        else:
            self._feature_processors = nn.ModuleDict(feature_processors)
        assert set(sum([config.feature_names for config in self._embedding_bag_collection.embedding_bag_configs()], [])) == set(
            feature_processors.keys() if not isinstance(feature_processors, FeatureProcessorsCollection) else sum(
                [config.feature_names for config in self._embedding_bag_collection.embedding_bag_configs()], [])), \
            "Passed in feature processors do not match feature names of embedding bag"
        assert embedding_bag_collection.is_weighted(), "EmbeddingBagCollection must accept weighted inputs for feature processor"

    def split(self) -> Tuple[FeatureProcessorsCollection, EmbeddingBagCollection]:
        if isinstance(self._feature_processors, nn.ModuleDict):
            return FeatureProcessorDictWrapper(self._feature_processors), self._embedding_bag_collection
        return self._feature_processors, self._embedding_bag_collection

    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        if isinstance(self._feature_processors, FeatureProcessorsCollection):
            fp_features = self._feature_processors(features)
        else:
            fp_features = apply_feature_processors_to_kjt(features, self._feature_processors)
        return self._embedding_bag_collection(fp_features)
