"""Session AUC (GAUC): the mean of the per-session AUCs.

Reference module: ``torchrec/metrics/gauc.py``: the examples of a batch arrive session by session, ``num_candidates`` gives the number
of examples of every session. A session contributes its weighted AUC unless all its labels are equal or all its predictions are equal;
additive states ``auc_sum`` / ``num_samples`` (the number of contributing sessions), lifetime + window; reported as ``gauc`` and
``gauc_num_samples``. (The group-by-key AUC over the sample window is ``AUCMetric(grouped_auc=True)``.)"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from ._bases import _SumStatesComputation
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401


def get_auc_states(labels: torch.Tensor, predictions: torch.Tensor, weights: torch.Tensor, num_candidates: torch.Tensor) -> Dict[str, torch.Tensor]:
    """[n_tasks, n] examples laid out session by session + the session lengths -> {auc_sum, num_samples} [n_tasks]."""
    n_tasks, n = predictions.shape
    lengths = num_candidates.reshape(-1).long()
    n_sess = int(lengths.numel())
    dev = predictions.device
    auc_sum = torch.zeros(n_tasks, dtype=torch.double, device=dev)
    count = torch.zeros(n_tasks, dtype=torch.double, device=dev)
    if n_sess == 0 or n == 0:
        return {"auc_sum": auc_sum, "num_samples": count}
    sess = torch.repeat_interleave(torch.arange(n_sess, device=dev), lengths, output_size=n)
    zeros = lambda: torch.zeros(n_sess, dtype=torch.double, device=dev)  # noqa: E731
    for t in range(n_tasks):
        p, y, w = predictions[t].double(), labels[t].double(), weights[t].double()
        by_p = torch.argsort(p, stable=True)  # ascending inside every session
        order = by_p[torch.argsort(sess[by_p], stable=True)]
        ys, ws, ss = y[order], w[order], sess[order]
        neg_w = ws * (1.0 - ys)
        cum_neg = torch.cumsum(neg_w, 0)
        base = zeros().index_add_(0, ss, neg_w)
        before_session = torch.cumsum(base, 0) - base  # negative weight of the earlier sessions
        num = zeros().index_add_(0, ss, ys * ws * (cum_neg - before_session[ss]))  # every positive: w_pos * (negative weight ranked at or below it)
        w_pos = zeros().index_add_(0, ss, ys * ws)
        auc = num / (w_pos * base + 1e-10)
        p_min = torch.full((n_sess,), float("inf"), dtype=torch.double, device=dev).scatter_reduce_(0, sess, p, reduce="amin")
        p_max = torch.full((n_sess,), float("-inf"), dtype=torch.double, device=dev).scatter_reduce_(0, sess, p, reduce="amax")
        keep = (w_pos > 0) & (base > 0) & (p_max > p_min)
        auc_sum[t] = (auc * keep).sum()
        count[t] = keep.sum()
    return {"auc_sum": auc_sum, "num_samples": count}


def compute_window_auc(auc: torch.Tensor, num_samples: torch.Tensor) -> Dict[str, torch.Tensor]:
    return {"gauc": (auc + 1e-9) / (num_samples + 2e-9), "num_samples": num_samples}


class GAUCMetricComputation(_SumStatesComputation):
    STATES = ["auc_sum", "num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        num_candidates = kwargs.get("num_candidates")
        if num_candidates is None and kwargs.get("required_inputs"):
            num_candidates = kwargs["required_inputs"].get("num_candidates")
        if predictions is None or weights is None:
            raise RecMetricException("Inputs 'predictions' and 'weights' should not be None for GAUCMetricComputation update")
        if num_candidates is None:
            raise RecMetricException("GAUCMetricComputation needs 'num_candidates' (the number of examples of every session)")
        return get_auc_states(labels, predictions, weights, num_candidates)

    def _reports(self, get, prefix):
        r = compute_window_auc(get("auc_sum"), get("num_samples"))
        return [MetricComputationReport(MetricName.GAUC, prefix, r["gauc"]), MetricComputationReport(MetricName.GAUC_NUM_SAMPLES, prefix, r["num_samples"])]


class GAUCMetric(RecMetric):
    _namespace: MetricNamespace = MetricNamespace.GAUC
    _computation_class = GAUCMetricComputation


def to_3d(tensor_2d: torch.Tensor, seq_lengths: torch.Tensor, max_length: int) -> torch.Tensor:
    """[n, n_task] rows laid out session by session -> [n_sessions, max_length, n_task], zero padded."""
    from ..ops import jagged as J

    return J.jagged_2d_to_dense(tensor_2d, J.asynchronous_complete_cumsum(seq_lengths.long()), max_length)


def compute_gauc_3d(predictions: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The padded form: [n_task, n_session, max_len] tensors (padding = zeros) -> {auc_sum, num_samples}; sessions with one label class
    or with identical (non-zero) predictions do not count."""
    order = torch.argsort(predictions, dim=-1)
    ys, ws = torch.gather(labels, -1, order), torch.gather(weights, -1, order)
    neg = ws * (1 - ys)
    num = (ys * ws * torch.cumsum(neg, dim=-1)).sum(-1)
    w_pos, w_neg = (ys * ws).sum(-1), neg.sum(-1)
    auc = num / (w_pos * w_neg + 1e-10)
    varied = ~torch.all((predictions == predictions[:, :, 0:1]) | (predictions == 0), dim=-1)
    keep = (w_pos > 0) & (w_neg > 0) & varied
    return {"auc_sum": (auc * keep).sum(-1), "num_samples": keep.sum(-1)}
