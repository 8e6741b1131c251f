"""NE per segment of a grouping key.

Reference module: ``torchrec/metrics/segmented_ne.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
from .ne import _ce, compute_ne  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


from dataclasses import dataclass
from typing import Union


@dataclass
class GroupingKeyConfig:
    """One segmentation: the name of the model output holding the group of every example, the number of groups, whether float keys
    are cast to int64."""

    name: str
    num_groups: int = 1
    cast_keys_to_int: bool = False


def _normalize_grouping_keys_config(grouping_keys, num_groups: int, cast_keys_to_int: bool) -> List[GroupingKeyConfig]:
    """``grouping_keys``: a name, a list of names, a list of dicts (name, [num_groups], [cast_keys_to_int]) or of GroupingKeyConfig."""
    if isinstance(grouping_keys, str):
        return [GroupingKeyConfig(grouping_keys, num_groups, cast_keys_to_int)]
    out = []
    for item in grouping_keys:
        if isinstance(item, str):
            out.append(GroupingKeyConfig(item, num_groups, cast_keys_to_int))
        elif isinstance(item, dict):
            out.append(GroupingKeyConfig(item["name"], item.get("num_groups", num_groups), item.get("cast_keys_to_int", cast_keys_to_int)))
        elif isinstance(item, GroupingKeyConfig):
            out.append(item)
        else:
            raise ValueError(f"Invalid grouping_keys item type: {type(item)}. Expected str, dict, or GroupingKeyConfig.")
    return out


_STATES = ["cross_entropy_sum", "weighted_num_samples", "pos_labels", "neg_labels"]


class SegmentedNEMetricComputation(RecMetricComputation):
    """NE per segment of the examples (binary labels): every grouping key is a model output in ``required_inputs`` with values in
    [0, num_groups). Reports ``segmented_ne`` (and ``logloss`` with ``include_logloss``) per group, lifetime + window; the description is
    ``_<group>`` for the single default key and ``_<group>@<key>`` otherwise."""

    def __init__(self, *args: Any, include_logloss: bool = False, num_groups: int = 1, grouping_keys: Union[str, List[Any]] = "grouping_keys",
                 cast_keys_to_int: bool = False, **kwargs: Any) -> None:
        self._include_logloss = include_logloss
        super().__init__(*args, **kwargs)
        self._grouping_key_configs = _normalize_grouping_keys_config(grouping_keys, num_groups, cast_keys_to_int)
        self._is_single_default_key = isinstance(grouping_keys, str) or (len(self._grouping_key_configs) == 1 and self._grouping_key_configs[0].name == "grouping_keys")
        self._num_groups = self._grouping_key_configs[0].num_groups
        self.eta = 1e-12
        for cfg in self._grouping_key_configs:
            for st in _STATES:
                self._add_state(self._state_prefix(cfg.name) + st, torch.zeros(self._n_tasks, cfg.num_groups, dtype=torch.double), add_window_state=True,
                                dist_reduce_fx="sum", persistent=True)

    def _state_prefix(self, key: str) -> str:
        return "" if self._is_single_default_key else f"{key}_"

    def _suffix(self, key: str) -> str:
        return "" if self._is_single_default_key else f"@{key}"

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        if predictions is None or weights is None:
            raise RecMetricException("Inputs 'predictions' and 'weights' should not be None for SegmentedNEMetricComputation update")
        req = kwargs.get("required_inputs")
        if req is None:
            legacy = kwargs.get("grouping_keys")  # keyword form of the single default key
            if legacy is None:
                raise RecMetricException(f"Required inputs for SegmentedNEMetricComputation update should be provided, got kwargs: {list(kwargs)}")
            req = {self._grouping_key_configs[0].name: legacy}
        n = predictions.shape[-1]
        ce = _ce(labels, predictions, weights, self.eta)
        w, y = weights.double(), labels.double()
        for cfg in self._grouping_key_configs:
            g = req.get(cfg.name)
            if g is None:
                raise RecMetricException(f"Required inputs for SegmentedNEMetricComputation update should contain '{cfg.name}', got keys: {list(req.keys())}")
            if g.dtype != torch.int64:
                if cfg.cast_keys_to_int and g.dtype in (torch.float32, torch.float64):
                    g = g.to(torch.int64)
                elif g.dtype in (torch.int32, torch.int16, torch.int8, torch.uint8):
                    g = g.to(torch.int64)
                else:
                    raise RecMetricException(f"Grouping key '{cfg.name}' expected to have type torch.int64 or torch.float32/torch.float64 with cast_keys_to_int set to true, got {g.dtype}.")
            idx = g.reshape(1, -1).expand(self._n_tasks, -1)
            for st, v in zip(_STATES, (ce, w, w * y, w * (1.0 - y))):
                name = self._state_prefix(cfg.name) + st
                state = getattr(self, name)
                delta = torch.zeros(self._n_tasks, cfg.num_groups, dtype=torch.double, device=state.device).scatter_add_(1, idx.to(state.device), v.to(state.device))
                state += delta
                self._aggregate_window_state(name, delta, n)

    def _compute(self) -> List[MetricComputationReport]:
        from .ne import compute_logloss

        reports: List[MetricComputationReport] = []
        for cfg in self._grouping_key_configs:
            pre, suf = self._state_prefix(cfg.name), self._suffix(cfg.name)
            scopes = [(MetricPrefix.LIFETIME, lambda n: getattr(self, n))]
            if self._batch_window_buffers is not None:
                scopes.append((MetricPrefix.WINDOW, lambda n: self.get_window_state(n)))
            nes = [(prefix, compute_ne(get(pre + "cross_entropy_sum"), get(pre + "weighted_num_samples"), get(pre + "pos_labels"), get(pre + "neg_labels"), self.eta)) for prefix, get in scopes]
            for gi in range(cfg.num_groups):
                for prefix, ne in nes:
                    reports.append(MetricComputationReport(MetricName.SEGMENTED_NE, prefix, ne[:, gi], description=f"_{gi}{suf}"))
            if self._include_logloss:
                lls = [(prefix, compute_logloss(get(pre + "cross_entropy_sum"), get(pre + "pos_labels"), get(pre + "neg_labels"), self.eta)) for prefix, get in scopes]
                for gi in range(cfg.num_groups):
                    for prefix, ll in lls:
                        reports.append(MetricComputationReport(MetricName.LOG_LOSS, prefix, ll[:, gi], description=f"_{gi}{suf}"))
        return reports


class SegmentedNEMetric(RecMetric):
    _namespace: MetricNamespace = MetricNamespace.SEGMENTED_NE
    _computation_class = SegmentedNEMetricComputation

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        for cfg in _normalize_grouping_keys_config(kwargs.get("grouping_keys", "grouping_keys"), kwargs.get("num_groups", 1), kwargs.get("cast_keys_to_int", False)):
            self._required_inputs.add(cfg.name)


# ---- stateless helpers of the reference module -----------------------------------------------------------------------------------------------------------------
from .ne import compute_cross_entropy, compute_logloss  # noqa: E402,F401


def get_segemented_ne_states_fused(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, grouping_keys: torch.Tensor, eta: float, num_groups: int,
                                   n_tasks: int) -> Dict[str, torch.Tensor]:
    """[n_tasks, n] inputs + one group id per example -> the four NE states as [n_tasks, num_groups] scatter sums."""
    idx = grouping_keys.long().reshape(1, -1).expand(n_tasks, -1)
    w, y = weights.double(), labels.double()

    def scatter(v: torch.Tensor) -> torch.Tensor:
        return torch.zeros(n_tasks, num_groups, dtype=torch.double, device=v.device).scatter_add_(1, idx, v.double())

    return {"cross_entropy_sum": scatter(compute_cross_entropy(labels, predictions, weights, eta)), "weighted_num_samples": scatter(w), "pos_labels": scatter(w * y),
            "neg_labels": scatter(w * (1.0 - y))}


def get_segemented_ne_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, grouping_keys: torch.Tensor, eta: float, num_groups: int) -> Dict[str, torch.Tensor]:
    """Single task form: states as [1, num_groups]."""
    return get_segemented_ne_states_fused(labels.reshape(1, -1), predictions.reshape(1, -1), weights.reshape(1, -1), grouping_keys, eta, num_groups, 1)


def compute_ne_helper(ce_sum: torch.Tensor, weighted_num_samples: torch.Tensor, pos_labels: torch.Tensor, neg_labels: torch.Tensor, eta: float) -> torch.Tensor:
    return compute_ne(ce_sum, weighted_num_samples, pos_labels, neg_labels, eta)


def compute_ne_fused(ce_sum: torch.Tensor, weighted_num_samples: torch.Tensor, pos_labels: torch.Tensor, neg_labels: torch.Tensor, num_groups: int, n_tasks: int,
                     eta: float) -> torch.Tensor:
    """NE of every (task, group) cell at once: [n_tasks, num_groups]."""
    return compute_ne(ce_sum.reshape(n_tasks, num_groups), weighted_num_samples.reshape(n_tasks, num_groups), pos_labels.reshape(n_tasks, num_groups),
                      neg_labels.reshape(n_tasks, num_groups), eta)
