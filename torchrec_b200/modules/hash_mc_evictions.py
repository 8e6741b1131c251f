"""Eviction scoring for the hash-based zero-collision table (MPZCH).

Reference: ``torchrec/modules/hash_mc_evictions.py`` - policy names :22, ``HashZchEvictionConfig`` :35, scorers :53-118 (single TTL, per-feature TTL), threshold /
opt-in eviction modules :121-213, ``HashZchEvictionModule`` :216. A slot's metadata is an "expire hour": ``now + ttl`` at (re)insertion; a slot may be
taken over by a colliding id once ``metadata < now``. The probe itself lives in ``hash_mc_modules.py`` (vectorised open addressing).
"""
from __future__ import annotations

import time
from enum import Enum, unique
from typing import Dict, List, Optional, Tuple

import torch

from ..sparse.jagged_tensor import JaggedTensor
from .hash_mc_modules import HashZchEvictionConfig, HashZchEvictionPolicyName  # noqa: F401


def get_kernel_from_policy(policy_name: Optional[HashZchEvictionPolicyName]) -> int:
    """0 = no eviction, 1 = threshold (TTL) eviction, 2 = LRU-by-metadata eviction - the selector the probe kernel takes."""
    if policy_name is None or policy_name == HashZchEvictionPolicyName.NONE:
        return 0
    if "TTL" in policy_name.name:
        return 1
    return 2


class HashZchEvictionScorer:
    def __init__(self, config: HashZchEvictionConfig) -> None:
        self._config = config

    def gen_score(self, feature: JaggedTensor, device: torch.device) -> torch.Tensor:
        return torch.empty(0, device=device)

    def gen_threshold(self) -> int:
        return -1


def _now_hours() -> int:
    return int(time.time() // 3600)


class HashZchSingleTtlScorer(HashZchEvictionScorer):
    """Every id of the module lives ``single_ttl`` hours after it was last seen."""

    def gen_score(self, feature: JaggedTensor, device: torch.device) -> torch.Tensor:
        assert self._config.single_ttl is not None and self._config.single_ttl > 0, "single_ttl must be set"
        return torch.full_like(feature.values(), self._config.single_ttl + _now_hours(), dtype=torch.int32, device=device)

    def gen_threshold(self) -> int:
        return _now_hours()


class HashZchPerFeatureTtlScorer(HashZchEvictionScorer):
    """TTL per feature: the scores follow the feature boundaries of the (flattened) input."""

    def __init__(self, config: HashZchEvictionConfig) -> None:
        super().__init__(config)
        ttls = getattr(config, "per_feature_ttl", None)
        assert ttls is not None and len(ttls) == len(config.features), "per_feature_ttl must have one entry per feature"
        self._per_feature_ttl = torch.tensor(ttls, dtype=torch.int32)

    def gen_score(self, feature: JaggedTensor, device: torch.device) -> torch.Tensor:
        F = self._per_feature_ttl.numel()
        lengths = feature.lengths().view(F, -1).sum(1)
        return torch.repeat_interleave(self._per_feature_ttl.to(device), lengths.to(device)).to(torch.int32) + _now_hours()

    def gen_threshold(self) -> int:
        return _now_hours()


def get_eviction_scorer(policy_name: HashZchEvictionPolicyName, config: HashZchEvictionConfig) -> HashZchEvictionScorer:
    if policy_name.name == "SINGLE_TTL_EVICTION":
        return HashZchSingleTtlScorer(config)
    if policy_name.name == "PER_FEATURE_TTL_EVICTION":
        return HashZchPerFeatureTtlScorer(config)
    return HashZchEvictionScorer(config)


class HashZchThresholdEvictionModule(torch.nn.Module):
    """Slots whose metadata dropped below the threshold are evictable (TTL expiry)."""

    def __init__(self, policy_name: HashZchEvictionPolicyName, config: HashZchEvictionConfig) -> None:
        super().__init__()
        self._policy_name = policy_name
        self._config = config
        self._scorer = get_eviction_scorer(policy_name, config)
        self._eviction_threshold = -1

    def extra_repr(self) -> str:
        return f"policy={self._policy_name.name}, features={self._config.features}"

    def build(self, feature: JaggedTensor, device: torch.device) -> Tuple[Optional[torch.Tensor], int]:
        self._eviction_threshold = self._scorer.gen_threshold()
        return self._scorer.gen_score(feature, device), self._eviction_threshold

    def evictable(self, metadata: torch.Tensor) -> torch.Tensor:
        return metadata.long() < self._eviction_threshold


class HashZchOptEvictionModule(torch.nn.Module):
    """Opt-in eviction: a colliding id takes the probed slot with the LOWEST metadata (least recently useful)."""

    def __init__(self, policy_name: HashZchEvictionPolicyName, config: HashZchEvictionConfig) -> None:
        super().__init__()
        self._policy_name, self._config = policy_name, config
        self._scorer = get_eviction_scorer(policy_name, config)

    def build(self, feature: JaggedTensor, device: torch.device) -> Tuple[Optional[torch.Tensor], int]:
        return self._scorer.gen_score(feature, device), -1

    def evictable(self, metadata: torch.Tensor) -> torch.Tensor:
        return torch.ones_like(metadata, dtype=torch.bool)


def get_eviction_module(policy_name: HashZchEvictionPolicyName, config: HashZchEvictionConfig) -> torch.nn.Module:
    return HashZchThresholdEvictionModule(policy_name, config) if get_kernel_from_policy(policy_name) == 1 else HashZchOptEvictionModule(policy_name, config)


class HashZchEvictionModule(torch.nn.Module):
    """One eviction sub-module per policy of a ZCH module; ``build`` returns the per-id scores and the threshold for this batch."""

    def __init__(self, policy_name: HashZchEvictionPolicyName, device: torch.device, config: HashZchEvictionConfig) -> None:
        super().__init__()
        self._policy_name = policy_name
        self._device = device
        self._eviction_module = get_eviction_module(policy_name, config)

    def forward(self, feature: JaggedTensor) -> Tuple[Optional[torch.Tensor], int]:
        return self._eviction_module.build(feature, self._device)

    def evictable(self, metadata: torch.Tensor) -> torch.Tensor:
        return self._eviction_module.evictable(metadata)
