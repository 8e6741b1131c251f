"""Regression AUC: fraction of correctly ordered (prediction, label) pairs.

Reference module: ``torchrec/metrics/rauc.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SampleBufferComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


class RAUCMetricComputation(_SampleBufferComputation):
    """Regression AUC: fraction of correctly ordered pairs (concordance) for continuous labels."""

    NAME = MetricName.RAUC

    def _value(self, p, l, w, extra):
        n = p.numel()
        if n < 2:
            return torch.tensor(0.5, dtype=torch.double)
        if n > 4096:  # subsample for the O(n^2) pair count
            idx = torch.randperm(n)[:4096]
            p, l = p[idx], l[idx]
        dp = p.unsqueeze(0) - p.unsqueeze(1)
        dl = l.unsqueeze(0) - l.unsqueeze(1)
        valid = dl != 0
        conc = ((dp * dl) > 0).double() + 0.5 * (dp == 0).double()
        return (conc * valid).sum() / (valid.sum() + EPS)


RAUCMetric = _make("RAUCMetric", RAUCMetricComputation, MetricNamespace.RAUC)


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
