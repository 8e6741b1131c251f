"""Session-level recall of the top-ranked items.

Reference module: ``torchrec/metrics/recall_session.py``: inside every update batch the examples are ranked within their session (the
session tensor is ``required_inputs[session_metric_def.session_var_name]``); the ``top_threshold`` best predictions of a session (ties
included) are the predicted positives; with ``run_ranking_of_labels`` the labels are ranked the same way. Additive weighted states
``num_true_pos`` / ``num_false_neg`` (lifetime + window); NaN until a positive example has been seen."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from ._bases import _SumStatesComputation
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecComputeMode, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401

NUM_TRUE_POS = "num_true_pos"
NUM_FALSE_NEGATIVE = "num_false_neg"
NUM_FALSE_POS = "num_false_pos"


def ranking_within_session(predictions: torch.Tensor, session: torch.Tensor) -> torch.Tensor:
    """[1, n] -> [n]: how many examples of the same session have a strictly larger value (0 = best, ties share a rank)."""
    p, s = predictions.reshape(-1), session.reshape(-1)
    _, sess = torch.unique(s, return_inverse=True)
    # sort by (session, value descending); the rank of an element = position of the first element of its tie group inside the session
    by_val = torch.argsort(p, descending=True, stable=True)
    order = by_val[torch.argsort(sess[by_val], stable=True)]
    ps, ss = p[order], sess[order]
    n = p.numel()
    pos = torch.arange(n, device=p.device)
    new_group = torch.ones(n, dtype=torch.bool, device=p.device)
    new_group[1:] = (ps[1:] != ps[:-1]) | (ss[1:] != ss[:-1])
    group_start = torch.cummax(torch.where(new_group, pos, torch.zeros_like(pos)), 0).values
    new_sess = torch.ones(n, dtype=torch.bool, device=p.device)
    new_sess[1:] = ss[1:] != ss[:-1]
    sess_start = torch.cummax(torch.where(new_sess, pos, torch.zeros_like(pos)), 0).values
    rank = torch.empty(n, dtype=torch.long, device=p.device)
    rank[order] = group_start - sess_start
    return rank


def _calc_num_true_pos(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return torch.sum(weights * labels * (predictions == 1).double(), dim=-1)


def _calc_num_false_neg(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return torch.sum(weights * labels * (predictions == 0).double(), dim=-1)


def _calc_num_false_pos(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    return torch.sum(weights * (1 - labels) * (predictions == 1).double(), dim=-1)


def _ratio_or_nan(num: torch.Tensor, other: torch.Tensor) -> torch.Tensor:
    d = num + other
    return torch.where(d == 0, torch.full_like(d, float("nan")), num / torch.where(d == 0, torch.ones_like(d), d))


class _SessionTopKComputation(_SumStatesComputation):
    """Shared part of session recall / precision: the session tensor, top-k labelling of predictions (and labels)."""

    def __init__(self, *args: Any, session_metric_def: Optional[Any] = None, top_threshold: Optional[int] = None, session_var_name: str = "session_ids",
                 run_ranking_of_labels: bool = False, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.top_threshold = getattr(session_metric_def, "top_threshold", None) if session_metric_def is not None else top_threshold
        if self.top_threshold is None:
            self.top_threshold = top_threshold if top_threshold is not None else 1
        self.run_ranking_of_labels = getattr(session_metric_def, "run_ranking_of_labels", run_ranking_of_labels)
        self.session_var_name = getattr(session_metric_def, "session_var_name", session_var_name)

    def _session(self, kwargs: Dict[str, Any]) -> torch.Tensor:
        req = kwargs.get("required_inputs") or {}
        session = req.get(self.session_var_name)
        if session is None:
            session = kwargs.get(self.session_var_name, kwargs.get("session_ids"))
        if session is None:
            raise RecMetricException(f"Need the {self.session_var_name} input to update the session metric")
        return session

    def _labelled(self, predictions, labels, weights, kwargs):
        session = self._session(kwargs)
        if predictions is None or weights is None:
            raise RecMetricException(f"Inputs 'predictions', 'weights' and 'session' should not be None for {type(self).__name__} update")
        session = session.reshape(1, -1)
        assert labels.dim() == 2 and labels.shape == predictions.shape == weights.shape == session.shape
        predictions, labels, weights = predictions.double(), labels.double(), weights.double()
        pred_pos = (ranking_within_session(predictions, session) < self.top_threshold).to(torch.int32).view(1, -1)
        if self.run_ranking_of_labels:
            labels = (ranking_within_session(labels, session) < self.top_threshold).to(torch.int32).view(1, -1).double()
        return labels, pred_pos, weights


class RecallSessionMetricComputation(_SessionTopKComputation):
    STATES = [NUM_TRUE_POS, NUM_FALSE_NEGATIVE]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        labels, pred_pos, weights = self._labelled(predictions, labels, weights, kwargs)
        return {NUM_TRUE_POS: _calc_num_true_pos(labels, pred_pos, weights), NUM_FALSE_NEGATIVE: _calc_num_false_neg(labels, pred_pos, weights)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.RECALL_SESSION_LEVEL, prefix, _ratio_or_nan(get(NUM_TRUE_POS), get(NUM_FALSE_NEGATIVE)))]


class _SessionRecMetric(RecMetric):
    """Session metrics are computed per task (no fused tasks / fused updates); every task needs a ``SessionMetricDef`` unless the
    top threshold / session tensor name are given as keyword arguments."""

    def __init__(self, world_size: int, my_rank: int, batch_size: int, tasks: List[Any], compute_mode: RecComputeMode = RecComputeMode.UNFUSED_TASKS_COMPUTATION,
                 window_size: int = 100, fused_update_limit: int = 0, **kwargs: Any) -> None:
        what = type(self).__name__
        if compute_mode in (RecComputeMode.FUSED_TASKS_COMPUTATION, RecComputeMode.FUSED_TASKS_AND_STATES_COMPUTATION):
            raise RecMetricException(f"Fused computation is not supported for session-level metrics ({what})")
        if fused_update_limit > 0:
            raise RecMetricException(f"Fused update is not supported for session-level metrics ({what})")
        for task in tasks:
            d = task.session_metric_def
            if d is None and "top_threshold" not in kwargs:
                raise RecMetricException("Please, specify the session metric definition")
            if d is not None and d.top_threshold is None:
                raise RecMetricException("Please, specify the top threshold")
        super().__init__(world_size=world_size, my_rank=my_rank, batch_size=batch_size, tasks=tasks, compute_mode=compute_mode, window_size=window_size,
                         fused_update_limit=fused_update_limit, **kwargs)
        for task in tasks:
            name = task.session_metric_def.session_var_name if task.session_metric_def is not None else kwargs.get("session_var_name", "session_ids")
            self._required_inputs.add(name)


class RecallSessionMetric(_SessionRecMetric):
    _namespace: MetricNamespace = MetricNamespace.RECALL_SESSION_LEVEL
    _computation_class = RecallSessionMetricComputation
