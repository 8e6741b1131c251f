"""Feature-score plumbing for virtual tables with score-based eviction (reference torchrec/distributed/feature_score_utils.py:24-182)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..modules.embedding_configs import BaseEmbeddingConfig, FeatureScoreBasedEvictionPolicy
from ..sparse.jagged_tensor import KeyedJaggedTensor


def create_sharding_type_to_feature_score_mapping(embedding_configs: Sequence[BaseEmbeddingConfig],
                                                   sharding_type_to_table_names: Dict[str, List[str]]) -> Tuple[bool, bool, Dict[str, Dict[str, float]]]:
    """Validate the virtual tables' eviction policies (score-based eviction is all-or-nothing across the virtual tables with an
    eviction policy) and build ``sharding type -> {feature: score weight}``. Returns (accumulate weights, auto collection, mapping)."""
    virtual = [c for c in embedding_configs if c.use_virtual_table and c.virtual_table_eviction_policy is not None]
    scored = [c for c in virtual if isinstance(c.virtual_table_eviction_policy, FeatureScoreBasedEvictionPolicy)]
    if not scored:
        return False, False, {}
    assert len(scored) == len(virtual), "if one virtual table uses FeatureScoreBasedEvictionPolicy, every virtual table with an eviction policy must"
    auto = [bool(c.virtual_table_eviction_policy.enable_auto_feature_score_collection) for c in scored]  # type: ignore[union-attr]
    assert all(auto) or not any(auto), "enable_auto_feature_score_collection must agree across virtual tables"
    by_name = {c.name: c for c in scored}
    mapping: Dict[str, Dict[str, float]] = {}
    for sharding_type, tables in sharding_type_to_table_names.items():
        fm: Dict[str, float] = {}
        for t in tables:
            c = by_name.get(t)
            if c is None:
                continue
            pol = c.virtual_table_eviction_policy
            assert isinstance(pol, FeatureScoreBasedEvictionPolicy)
            for f in c.feature_names:
                if pol.feature_score_mapping and f in pol.feature_score_mapping:
                    fm[f] = float(pol.feature_score_mapping[f])
                elif pol.feature_score_default_value is not None:
                    fm[f] = float(pol.feature_score_default_value)
                else:
                    assert auto[0], f"feature {f} of virtual table {t} has no score and auto collection is off"
                    fm[f] = 0.0
        if fm:
            mapping[sharding_type] = fm
    return True, auto[0], mapping


def may_collect_feature_scores(features: KeyedJaggedTensor, enabled: bool, feature_score_mapping: Dict[str, float]) -> KeyedJaggedTensor:
    """Attach per-id scores as KJT weights (``weight[i] = score of the feature id i belongs to``) so the key-value cache can rank rows."""
    if not enabled or not feature_score_mapping:
        return features
    lpk = features.length_per_key()
    w = torch.cat([torch.full((n,), float(feature_score_mapping.get(k, 0.0)), dtype=torch.float32, device=features.device()) for k, n in zip(features.keys(), lpk)]) \
        if lpk else torch.zeros(0, dtype=torch.float32, device=features.device())
    return KeyedJaggedTensor(keys=features.keys(), values=features.values(), weights=w, lengths=features.lengths(), stride=features.stride(), length_per_key=lpk)
