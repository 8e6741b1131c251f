"""Session-level normalised discounted cumulative gain (NDCG @ k).

Reference module: ``torchrec/metrics/ndcg.py``: additive states ``sum_ndcg`` / ``num_sessions`` (lifetime + window), sessions are formed
inside every update batch, the report is ``1 - NDCG`` by default (a decreasing "loss" curve next to NE), scaled by the largest weight of
the session unless the gains themselves are scaled by the weights; a session without any gain counts with NDCG 0. The session ids arrive
as ``required_inputs[session_key]`` (or as the keyword ``session_ids``)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from ._bases import _SumStatesComputation
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401

SUM_NDCG = "sum_ndcg"
NUM_SESSIONS = "num_sessions"
REQUIRED_INPUTS = "required_inputs"
SESSION_KEY = "session_id"


def _rank_in_session(scores: torch.Tensor, sess: torch.Tensor, n_sessions: int):
    """Order of the samples by (session, descending score) and the 1-based rank of every sample inside its session."""
    by_score = torch.argsort(scores, descending=True, stable=True)
    order = by_score[torch.argsort(sess[by_score], stable=True)]
    counts = torch.bincount(sess, minlength=n_sessions)
    start = torch.cumsum(counts, 0) - counts
    rank = torch.arange(scores.numel(), device=scores.device) - start[sess[order]] + 1
    return order, rank


def _get_ndcg_states(*, labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, session_ids: torch.Tensor, exponential_gain: bool, k: int = -1,
                     report_ndcg_as_decreasing_curve: bool = True, remove_single_length_sessions: bool = False,
                     scale_by_weights_tensor: bool = False) -> Dict[str, torch.Tensor]:
    """[n_tasks, n] inputs (session ids identical for every task) -> {sum_ndcg [n_tasks], num_sessions [n_tasks]} of this batch."""
    sid = session_ids[0] if session_ids.dim() == 2 else session_ids
    _, sess, lengths = torch.unique(sid, return_inverse=True, return_counts=True)
    if remove_single_length_sessions:
        keep = lengths[sess] > 1
        predictions, labels, weights = predictions[:, keep], labels[:, keep], weights[:, keep]
        _, sess, lengths = torch.unique(sess[keep], return_inverse=True, return_counts=True)
    n_sessions = int(lengths.numel())
    if n_sessions == 0:
        return {}
    if scale_by_weights_tensor:
        labels, predictions = weights * labels, weights * predictions
    top = int(lengths.max()) if k == -1 else min(int(k), int(lengths.max()))
    n_tasks = labels.shape[0]
    out = torch.zeros(n_tasks, dtype=torch.double, device=labels.device)
    for t in range(n_tasks):
        gains = torch.exp2(labels[t]) - 1.0 if exponential_gain else labels[t]
        dcgs = []
        for scores in (predictions[t], labels[t]):  # observed order, ideal order
            order, rank = _rank_in_session(scores, sess, n_sessions)
            disc = torch.where(rank <= top, torch.reciprocal(torch.log2(rank.double() + 1.0)), torch.zeros((), dtype=torch.double, device=labels.device))
            dcgs.append(torch.zeros(n_sessions, dtype=torch.double, device=labels.device).index_add_(0, sess[order], gains[order].double() * disc))
        observed, ideal = dcgs
        ideal = torch.where(ideal == 0, torch.full_like(ideal, 1e-6), ideal)
        ndcg = observed / ideal
        report = (1.0 - ndcg) if report_ndcg_as_decreasing_curve else ndcg
        if not scale_by_weights_tensor:
            max_w = torch.zeros(n_sessions, dtype=torch.double, device=labels.device).scatter_reduce_(0, sess, weights[t].double(), reduce="amax", include_self=True)
            report = report * max_w
        out[t] = report.sum()
    return {SUM_NDCG: out, NUM_SESSIONS: torch.full((n_tasks,), float(n_sessions), dtype=torch.double, device=labels.device)}


def _compute_ndcg(*, sum_ndcg: torch.Tensor, num_sessions: torch.Tensor) -> torch.Tensor:
    return sum_ndcg / num_sessions


class NDCGComputation(_SumStatesComputation):
    """NDCG @ k over the sessions of every batch (see the module docstring for what is reported)."""

    STATES = [SUM_NDCG, NUM_SESSIONS]

    def __init__(self, *args: Any, exponential_gain: bool = False, session_key: str = SESSION_KEY, k: int = -1, report_ndcg_as_decreasing_curve: bool = True,
                 remove_single_length_sessions: bool = False, scale_by_weights_tensor: bool = False, is_negative_task_mask: Optional[List[bool]] = None,
                 **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._exponential_gain = exponential_gain
        self._session_key = session_key
        self._k = k
        self._report_ndcg_as_decreasing_curve = report_ndcg_as_decreasing_curve
        self._remove_single_length_sessions = remove_single_length_sessions
        self._scale_by_weights_tensor = scale_by_weights_tensor
        self._is_negative_task_mask = is_negative_task_mask

    def _batch_states(self, predictions, labels, weights, **kwargs):
        session_ids = None
        if REQUIRED_INPUTS in kwargs and kwargs[REQUIRED_INPUTS] is not None:
            session_ids = kwargs[REQUIRED_INPUTS].get(self._session_key)
        if session_ids is None:
            session_ids = kwargs.get("session_ids", kwargs.get(self._session_key))
        if session_ids is None:
            raise RecMetricException(f"session key {self._session_key!r} should be in the required inputs: it is needed to calculate the NDCG loss")
        if predictions is None or weights is None:
            raise RecMetricException("Inputs 'predictions', 'weights' and 'session_ids' should not be None for NDCGMetricComputation update")
        predictions, labels, weights = predictions.double(), labels.double(), weights.double()
        if self._is_negative_task_mask is not None:  # e.g. p(skip): a label of 0 should rank first
            mask = torch.tensor(self._is_negative_task_mask, dtype=torch.bool, device=predictions.device)
            predictions = torch.where(mask[:, None], 1.0 - predictions, predictions)
            labels = torch.where(mask[:, None], 1.0 - labels, labels)
        session_ids = session_ids.reshape(1, -1).expand(labels.shape[0], -1) if session_ids.dim() == 1 else session_ids
        assert predictions.shape == labels.shape == weights.shape == session_ids.shape and predictions.dim() == 2 and predictions.numel() > 0
        assert bool((session_ids[0] == session_ids).all()), "every task must see the same session ids"
        return _get_ndcg_states(labels=labels, predictions=predictions, weights=weights, session_ids=session_ids, exponential_gain=self._exponential_gain, k=self._k,
                                report_ndcg_as_decreasing_curve=self._report_ndcg_as_decreasing_curve,
                                remove_single_length_sessions=self._remove_single_length_sessions, scale_by_weights_tensor=self._scale_by_weights_tensor)

    def _reports(self, get, prefix):
        s, n = get(SUM_NDCG), get(NUM_SESSIONS)
        return [MetricComputationReport(MetricName.NDCG, prefix, torch.where(n > 0, s / n.clamp(min=1.0), torch.zeros_like(s)))]


NDCGMetricComputation = NDCGComputation


class NDCGMetric(RecMetric):
    _namespace: MetricNamespace = MetricNamespace.NDCG
    _computation_class = NDCGComputation

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._required_inputs.add(kwargs.get("session_key", SESSION_KEY))
