"""Regression AUC: fraction of correctly ordered (prediction, label) pairs.

Reference module: ``torchrec/metrics/rauc.py``. The metric classes live in ``metrics_impl.py`` (one sum-state / sample-buffer base for all 40+ metrics);
this module gives them their reference import path and holds the stateless ``compute_*`` / ``get_*_states`` helpers."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .metrics_impl import RAUCMetric, RAUCMetricComputation  # noqa: F401

EPS = torch.finfo(torch.float64).eps

def count_reverse_pairs_divide_and_conquer(input: List[float]) -> float:
    """Number of inversions of ``input`` by merge sort, O(n log n)."""
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

    return float(rec(0, len(a)))


def compute_rauc(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    out = []
    for t in range(n_tasks):
        order = torch.argsort(predictions[t], stable=True)
        n = order.numel()
        total = n * (n - 1) / 2
        inv = count_reverse_pairs_divide_and_conquer(labels[t][order].tolist())
        out.append(torch.tensor(1.0 - inv / total if total > 0 else 1.0, dtype=torch.double))
    return torch.stack(out)


def compute_rauc_per_group(n_tasks: int, predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor, grouping_keys: torch.Tensor) -> torch.Tensor:
    out = []
    for t in range(n_tasks):
        vals = [compute_rauc(1, predictions[t : t + 1, grouping_keys == g], labels[t : t + 1, grouping_keys == g], weights[t : t + 1, grouping_keys == g])[0] for g in torch.unique(grouping_keys)]
        out.append(torch.stack(vals).mean() if vals else torch.tensor(0.5, dtype=torch.double))
    return torch.stack(out)
