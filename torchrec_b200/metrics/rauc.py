"""Regression AUC: fraction of correctly ordered (prediction, label) pairs.

Reference module: ``torchrec/metrics/rauc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


GROUPING_KEYS = "grouping_keys"


class RAUCMetricComputation(_SampleBufferComputation):
    """Regression AUC over the sample window: 1 - (pairs ordered the wrong way round by the prediction) / (all pairs), counted by
    merge sort. ``grouped_rauc`` adds the mean over the groups of ``required_inputs['grouping_keys']`` (``grouped_rauc``)."""

    NAME = MetricName.RAUC

    def __init__(self, *args: Any, grouped_rauc: bool = False, fused_update_limit: int = 0, **kwargs: Any) -> None:
        if grouped_rauc and fused_update_limit > 0:
            raise RecMetricException("Grouped RAUC and Fused Update Limit cannot be enabled together yet.")
        self._grouped_rauc = grouped_rauc
        self.EXTRA = [GROUPING_KEYS] if grouped_rauc else []
        super().__init__(*args, fused_update_limit=fused_update_limit, **kwargs)

    def _value(self, p, l, w, extra):
        return compute_rauc(1, p.view(1, -1), l.view(1, -1), w.view(1, -1))[0]

    def _compute(self) -> List[MetricComputationReport]:
        reports = super()._compute()
        if self._grouped_rauc:
            reports.append(MetricComputationReport(MetricName.GROUPED_RAUC, MetricPrefix.WINDOW,
                                                   compute_rauc_per_group(self._n_tasks, self.predictions, self.labels, self.weights, getattr(self, GROUPING_KEYS)[0])))
        return reports


class RAUCMetric(RecMetric):
    _namespace: MetricNamespace = MetricNamespace.RAUC
    _computation_class = RAUCMetricComputation

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if kwargs.get("grouped_rauc"):
            self._required_inputs.add(GROUPING_KEYS)


def count_reverse_pairs_divide_and_conquer(input: List[float]) -> float:
    """Fraction of the pairs of ``input`` that are inverted (i < j, input[i] > input[j]), counted by merge sort in O(n log n): 0 for an
    ascending list, 1 for a strictly descending one (0 for fewer than two elements)."""
    a = list(input)

    def rec(lo: int, hi: int) -> int:
        if hi - lo <= 1:
            return 0
        mid = (lo + hi) // 2
        n = rec(lo, mid) + rec(mid, hi)
        merged, i, j = [], lo, mid
        while i < mid and j < hi:
            if a[i] <= a[j]:
                merged.append(a[i]); i += 1
            else:
                merged.append(a[j]); j += 1; n += mid - i
        merged.extend(a[i:mid]); merged.extend(a[j:hi])
        a[lo:hi] = merged
        return n

    n = len(a)
    return float(rec(0, n)) / (n * (n - 1) / 2) if n > 1 else 0.0


def compute_rauc(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    out = []
    for t in range(n_tasks):
        by_label = torch.argsort(labels[t], stable=True)  # ties of the prediction are ordered by label: they count as ordered correctly
        order = by_label[torch.argsort(predictions[t][by_label], stable=True)]
        out.append(torch.tensor(1.0 - count_reverse_pairs_divide_and_conquer(labels[t][order].tolist()), dtype=torch.double))
    return torch.stack(out)


def compute_rauc_per_group(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, grouping_keys: torch.Tensor) -> torch.Tensor:
    out = []
    for t in range(n_tasks):
        vals = [compute_rauc(1, predictions[t : t + 1, grouping_keys == g], labels[t : t + 1, grouping_keys == g], weights[t : t + 1, grouping_keys == g])[0] for g in torch.unique(grouping_keys)]
        out.append(torch.stack(vals).mean() if vals else torch.tensor(0.5, dtype=torch.double))
    return torch.stack(out)


def conquer_and_count(input: List[float], left_index: int, mid_index: int, right_index: int) -> int:
    """Merge the sorted halves ``input[left:mid+1]`` and ``input[mid+1:right+1]`` in place; returns the inversions between them."""
    left, right = input[left_index : mid_index + 1], input[mid_index + 1 : right_index + 1]
    i = j = inversions = 0
    k = left_index
    while i < len(left) and j < len(right):
        if left[i] <= right[j]:
            input[k] = left[i]
            i += 1
        else:
            input[k] = right[j]
            j += 1
            inversions += len(left) - i
        k += 1
    input[k : right_index + 1] = left[i:] + right[j:]
    return inversions


def divide(input: List[float], low: int, high: int) -> int:
    """Merge sort of ``input[low:high+1]`` in place; returns its number of inversions."""
    if low >= high:
        return 0
    mid = low + (high - low) // 2
    return divide(input, low, mid) + divide(input, mid + 1, high) + conquer_and_count(input, low, mid, high)
