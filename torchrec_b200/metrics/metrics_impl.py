"""Metric implementations on the RecMetric framework (reference torchrec/metrics/*.py — one file per metric there).

Additive-state metrics: NE (+logloss, positive NE), Calibration, CTR, MSE/RMSE, MAE, Accuracy, Precision, Recall,
WeightedAvg, TensorWeightedAvg, Average, Scalar, Output, TowerQPS, MulticlassRecall, NMSE, UnweightedNE, CaliFreeNE,
ServingNE, ServingCalibration, HindsightTargetPR. Sample-buffer metrics: AUC, AUPRC, RAUC, GroupedAUC (GAUC), XAUC, NDCG,
Recall/Precision at session level, SegmentedNE, MultiLabelPrecision."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Type

import torch

from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException

EPS = torch.finfo(torch.float64).eps


def _zeros(n: int) -> torch.Tensor:
    return torch.zeros(n, dtype=torch.double)


class _SumStatesComputation(RecMetricComputation):
    """Base for metrics whose states are per-task sums: subclasses define STATES and ``_batch_states`` / ``_value``."""

    STATES: List[str] = []

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        for s in self.STATES:
            self._add_state(s, _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="sum", persistent=True)

    def _batch_states(self, predictions, labels, weights, **kwargs) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    def _reports(self, get) -> List[MetricComputationReport]:
        raise NotImplementedError

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        if predictions is None and "predictions" in self._needs():
            raise RecMetricException(f"Inputs 'predictions' should not be None for {type(self).__name__} update")
        states = self._batch_states(predictions, labels, weights, **kwargs)
        n = labels.shape[-1]
        for name, v in states.items():
            st = getattr(self, name)
            v = v.to(st.dtype).to(st.device)
            st += v
            self._aggregate_window_state(name, v, n)

    def _needs(self) -> List[str]:
        return ["predictions"]

    def _compute(self) -> List[MetricComputationReport]:
        reports = self._reports(lambda n: getattr(self, n), MetricPrefix.LIFETIME)
        if self._batch_window_buffers is not None:
            reports += self._reports(lambda n: self.get_window_state(n), MetricPrefix.WINDOW)
        return reports


def _ce(labels, preds, weights, eta=1e-12):
    p = torch.clamp(preds.double(), eta, 1 - eta)
    return -weights.double() * (labels.double() * torch.log2(p) + (1 - labels.double()) * torch.log2(1 - p))


def compute_ne(ce_sum, weighted_num_samples, pos_labels, neg_labels, eta=1e-12) -> torch.Tensor:
    mean_label = pos_labels / (weighted_num_samples + EPS)
    ce_norm = -(pos_labels * torch.log2(mean_label + eta) + neg_labels * torch.log2(1 - mean_label + eta))
    return ce_sum / (ce_norm + EPS)


class NEMetricComputation(_SumStatesComputation):
    """Normalized entropy = cross entropy / entropy of the base rate. ``include_logloss`` adds logloss."""

    STATES = ["cross_entropy_sum", "weighted_num_samples", "pos_labels", "neg_labels"]

    def __init__(self, *args: Any, include_logloss: bool = False, allow_missing_label_with_zero_weight: bool = False, **kwargs: Any) -> None:
        self._include_logloss = include_logloss
        super().__init__(*args, **kwargs)
        self.eta = 1e-12

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"cross_entropy_sum": _ce(labels, predictions, weights, self.eta).sum(-1), "weighted_num_samples": weights.double().sum(-1),
                "pos_labels": (weights.double() * labels.double()).sum(-1), "neg_labels": (weights.double() * (1 - labels.double())).sum(-1)}

    def _reports(self, get, prefix):
        ne = compute_ne(get("cross_entropy_sum"), get("weighted_num_samples"), get("pos_labels"), get("neg_labels"), self.eta)
        out = [MetricComputationReport(MetricName.NE, prefix, ne)]
        if self._include_logloss:
            ll = get("cross_entropy_sum") / (get("weighted_num_samples") + EPS) * torch.log(torch.tensor(2.0, dtype=torch.double))
            out.append(MetricComputationReport(MetricName.LOG_LOSS, prefix, ll))
        return out


class CalibrationMetricComputation(_SumStatesComputation):
    STATES = ["calibration_num", "calibration_denom"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"calibration_num": (predictions.double() * weights.double()).sum(-1), "calibration_denom": (labels.double() * weights.double()).sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.CALIBRATION, prefix, get("calibration_num") / (get("calibration_denom") + EPS))]


class CTRMetricComputation(_SumStatesComputation):
    STATES = ["ctr_num", "ctr_denom"]

    def _needs(self):
        return []

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"ctr_num": (labels.double() * weights.double()).sum(-1), "ctr_denom": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.CTR, prefix, get("ctr_num") / (get("ctr_denom") + EPS))]


class MSEMetricComputation(_SumStatesComputation):
    STATES = ["error_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        d = predictions.double() - labels.double()
        return {"error_sum": (weights.double() * d * d).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        mse = get("error_sum") / (get("weighted_num_samples") + EPS)
        return [MetricComputationReport(MetricName.MSE, prefix, mse), MetricComputationReport(MetricName.RMSE, prefix, torch.sqrt(mse))]


class MAEMetricComputation(_SumStatesComputation):
    STATES = ["error_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"error_sum": (weights.double() * (predictions.double() - labels.double()).abs()).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.MAE, prefix, get("error_sum") / (get("weighted_num_samples") + EPS))]


class NMSEMetricComputation(_SumStatesComputation):
    """MSE normalised by the MSE of the constant predictor 1."""

    STATES = ["error_sum", "const_pred_error_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        d = predictions.double() - labels.double()
        c = 1.0 - labels.double()
        return {"error_sum": (weights.double() * d * d).sum(-1), "const_pred_error_sum": (weights.double() * c * c).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        nmse = get("error_sum") / (get("const_pred_error_sum") + EPS)
        return [MetricComputationReport(MetricName.NMSE, prefix, nmse), MetricComputationReport(MetricName.NRMSE, prefix, torch.sqrt(nmse))]


class AccuracyMetricComputation(_SumStatesComputation):
    STATES = ["accuracy_sum", "weighted_num_samples"]

    def __init__(self, *args: Any, threshold: float = 0.5, **kwargs: Any) -> None:
        self._threshold = threshold
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        pred = (predictions.double() >= self._threshold).double()
        return {"accuracy_sum": (weights.double() * (pred == labels.double()).double()).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.ACCURACY, prefix, get("accuracy_sum") / (get("weighted_num_samples") + EPS))]


class PrecisionMetricComputation(_SumStatesComputation):
    STATES = ["true_pos_sum", "false_pos_sum"]

    def __init__(self, *args: Any, threshold: float = 0.5, **kwargs: Any) -> None:
        self._threshold = threshold
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        pred = (predictions.double() >= self._threshold).double()
        return {"true_pos_sum": (weights.double() * pred * labels.double()).sum(-1), "false_pos_sum": (weights.double() * pred * (1 - labels.double())).sum(-1)}

    def _reports(self, get, prefix):
        tp, fp = get("true_pos_sum"), get("false_pos_sum")
        return [MetricComputationReport(MetricName.PRECISION, prefix, torch.where(tp + fp == 0.0, torch.zeros_like(tp), tp / (tp + fp)))]


class RecallMetricComputation(_SumStatesComputation):
    STATES = ["true_pos_sum", "false_neg_sum"]

    def __init__(self, *args: Any, threshold: float = 0.5, **kwargs: Any) -> None:
        self._threshold = threshold
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        pred = (predictions.double() >= self._threshold).double()
        return {"true_pos_sum": (weights.double() * pred * labels.double()).sum(-1), "false_neg_sum": (weights.double() * (1 - pred) * labels.double()).sum(-1)}

    def _reports(self, get, prefix):
        tp, fn = get("true_pos_sum"), get("false_neg_sum")
        return [MetricComputationReport(MetricName.RECALL, prefix, torch.where(tp + fn == 0.0, torch.zeros_like(tp), tp / (tp + fn)))]


class WeightedAvgMetricComputation(_SumStatesComputation):
    STATES = ["weighted_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"weighted_sum": (predictions.double() * weights.double()).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.WEIGHTED_AVG, prefix, get("weighted_sum") / (get("weighted_num_samples") + EPS))]


class AverageMetricComputation(_SumStatesComputation):
    """Mean of the labels (useful for tracking target statistics)."""

    STATES = ["sum", "num_samples"]

    def _needs(self):
        return []

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"sum": (labels.double() * weights.double()).sum(-1), "num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.AVERAGE, prefix, get("sum") / (get("num_samples") + EPS))]


class ScalarMetricComputation(RecMetricComputation):
    """Reports the last observed scalar (labels carry the value) and its window average."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._add_state("labels", _zeros(self._n_tasks), add_window_state=False, dist_reduce_fx="max", persistent=False)
        self._add_state("window_count", _zeros(self._n_tasks), add_window_state=False, dist_reduce_fx="sum", persistent=False)
        self._add_state("window_sum", _zeros(self._n_tasks), add_window_state=False, dist_reduce_fx="sum", persistent=False)

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        self.labels = labels.double().mean(-1).to(self.labels.device)
        self.window_count += 1
        self.window_sum += labels.double().mean(-1).to(self.window_sum.device)

    def _compute(self) -> List[MetricComputationReport]:
        return [MetricComputationReport(MetricName.SCALAR, MetricPrefix.LIFETIME, self.labels),
                MetricComputationReport(MetricName.SCALAR, MetricPrefix.WINDOW, self.window_sum / (self.window_count + EPS))]


class TowerQPSMetricComputation(RecMetricComputation):
    """Examples per second seen by a tower (lifetime and window), max over ranks of the elapsed time."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self._warmup_steps = kwargs.pop("warmup_steps", 0)
        super().__init__(*args, **kwargs)
        self._add_state("num_examples", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="sum", persistent=True)
        self._add_state("time_lapse", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="max", persistent=True)
        self._steps = 0
        self._previous_ts = 0.0

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        self._steps += 1
        if self._steps <= self._warmup_steps:
            return
        ts = time.monotonic()
        if self._steps == self._warmup_steps + 1:
            self._previous_ts = ts
            return
        n = torch.full((self._n_tasks,), float(labels.shape[-1]), dtype=torch.double)
        dt = torch.full((self._n_tasks,), ts - self._previous_ts, dtype=torch.double)
        self.num_examples += n.to(self.num_examples.device)
        self.time_lapse += dt.to(self.time_lapse.device)
        self._aggregate_window_state("num_examples", n, labels.shape[-1])
        self._aggregate_window_state("time_lapse", dt, labels.shape[-1])
        self._previous_ts = ts

    def _compute(self) -> List[MetricComputationReport]:
        out = [MetricComputationReport(MetricName.TOWER_QPS, MetricPrefix.LIFETIME, self.num_examples / (self.time_lapse + EPS))]
        if self._batch_window_buffers is not None:
            out.append(MetricComputationReport(MetricName.TOWER_QPS, MetricPrefix.WINDOW, self.get_window_state("num_examples") / (self.get_window_state("time_lapse") + EPS)))
        return out


class MulticlassRecallMetricComputation(RecMetricComputation):
    """recall@k for k in 1..number_of_classes; predictions [n_tasks, B, C]."""

    def __init__(self, *args: Any, number_of_classes: int = 2, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._number_of_classes = number_of_classes
        self._add_state("tp_at_k", torch.zeros(self._n_tasks, number_of_classes, dtype=torch.double), add_window_state=True, dist_reduce_fx="sum")
        self._add_state("total_weights", _zeros(self._n_tasks), add_window_state=True, dist_reduce_fx="sum")

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        C = self._number_of_classes
        p = predictions.reshape(self._n_tasks, -1, C)
        ranks = torch.argsort(p, dim=-1, descending=True)
        l = labels.reshape(self._n_tasks, -1).long()
        w = weights.reshape(self._n_tasks, -1).double()
        hit_pos = (ranks == l.unsqueeze(-1)).double().argmax(-1)  # rank position of the true class
        tp = torch.zeros(self._n_tasks, C, dtype=torch.double)
        for k in range(C):
            tp[:, k] = (w * (hit_pos <= k).double()).sum(-1)
        self.tp_at_k += tp.to(self.tp_at_k.device)
        self.total_weights += w.sum(-1).to(self.total_weights.device)
        self._aggregate_window_state("tp_at_k", tp, l.shape[-1])
        self._aggregate_window_state("total_weights", w.sum(-1), l.shape[-1])

    def _compute(self) -> List[MetricComputationReport]:
        out = [MetricComputationReport(MetricName.MULTICLASS_RECALL, MetricPrefix.LIFETIME, self.tp_at_k / (self.total_weights.unsqueeze(-1) + EPS))]
        if self._batch_window_buffers is not None:
            out.append(MetricComputationReport(MetricName.MULTICLASS_RECALL, MetricPrefix.WINDOW,
                                               self.get_window_state("tp_at_k") / (self.get_window_state("total_weights").unsqueeze(-1) + EPS)))
        return out


# ---- sample-buffer metrics --------------------------------------------------------------------------------------
def _auc_from_samples(preds: torch.Tensor, labels: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """Weighted ROC AUC with tie handling (trapezoid over the sorted-by-score cumulative TP/FP curve)."""
    if preds.numel() == 0:
        return torch.tensor(0.5, dtype=torch.double)
    order = torch.argsort(preds, descending=True)
    p, l, w = preds[order].double(), labels[order].double(), weights[order].double()
    ctp = torch.cumsum(w * l, 0)
    cfp = torch.cumsum(w * (1 - l), 0)
    # keep only the last point of every group of tied scores
    distinct = torch.ones_like(p, dtype=torch.bool)
    distinct[:-1] = p[1:] != p[:-1]
    ctp, cfp = ctp[distinct], cfp[distinct]
    ctp = torch.cat([ctp.new_zeros(1), ctp])
    cfp = torch.cat([cfp.new_zeros(1), cfp])
    if ctp[-1] == 0 or cfp[-1] == 0:
        return torch.tensor(0.5, dtype=torch.double)
    return torch.trapz(ctp, cfp) / (ctp[-1] * cfp[-1])


def _auprc_from_samples(preds, labels, weights) -> torch.Tensor:
    if preds.numel() == 0:
        return torch.tensor(0.0, dtype=torch.double)
    order = torch.argsort(preds, descending=True)
    p, l, w = preds[order].double(), labels[order].double(), weights[order].double()
    ctp = torch.cumsum(w * l, 0)
    cfp = torch.cumsum(w * (1 - l), 0)
    distinct = torch.ones_like(p, dtype=torch.bool)
    distinct[:-1] = p[1:] != p[:-1]
    ctp, cfp = ctp[distinct], cfp[distinct]
    if ctp[-1] == 0:
        return torch.tensor(0.0, dtype=torch.double)
    precision = ctp / (ctp + cfp + EPS)
    recall = ctp / ctp[-1]
    recall_prev = torch.cat([recall.new_zeros(1), recall[:-1]])
    return ((recall - recall_prev) * precision).sum()


class _SampleBufferComputation(RecMetricComputation):
    """Keeps the last ``window_size`` (pred, label, weight[, group]) samples per task; all-gathered at compute."""

    EXTRA: List[str] = []

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        for n in ["predictions", "labels", "weights"] + self.EXTRA:
            self._add_state(n, torch.zeros(self._n_tasks, 0, dtype=torch.double), add_window_state=False, dist_reduce_fx="cat", persistent=False)

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        if predictions is None or weights is None:
            raise RecMetricException(f"Inputs 'predictions' and 'weights' should not be None for {type(self).__name__} update")
        vals = {"predictions": predictions, "labels": labels, "weights": weights}
        for e in self.EXTRA:
            v = kwargs.get(e)
            if v is None and "required_inputs" in kwargs:
                v = kwargs["required_inputs"].get(e)
            if v is None:
                raise RecMetricException(f"{type(self).__name__} needs '{e}'")
            vals[e] = v.reshape(1, -1).expand(self._n_tasks, -1) if v.dim() == 1 else v
        cap = self._window_size
        for n, v in vals.items():
            cur = getattr(self, n)
            new = torch.cat([cur, v.reshape(self._n_tasks, -1).double().to(cur.device)], dim=-1)
            if new.shape[-1] > cap:
                new = new[:, -cap:]
            self._buffers[n] = new

    def _value(self, p, l, w, extra) -> torch.Tensor:
        raise NotImplementedError

    NAME = MetricName.AUC

    def _compute(self) -> List[MetricComputationReport]:
        vals = []
        for t in range(self._n_tasks):
            extra = {e: getattr(self, e)[t] for e in self.EXTRA}
            vals.append(self._value(self.predictions[t], self.labels[t], self.weights[t], extra))
        return [MetricComputationReport(self.NAME, MetricPrefix.WINDOW, torch.stack(vals))]

    def reset(self) -> None:
        for n in ["predictions", "labels", "weights"] + self.EXTRA:
            self._buffers[n] = torch.zeros(self._n_tasks, 0, dtype=torch.double)


class AUCMetricComputation(_SampleBufferComputation):
    NAME = MetricName.AUC

    def _value(self, p, l, w, extra):
        return _auc_from_samples(p, l, w)


class AUPRCMetricComputation(_SampleBufferComputation):
    NAME = MetricName.AUPRC

    def _value(self, p, l, w, extra):
        return _auprc_from_samples(p, l, w)


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


class GroupedAUCMetricComputation(_SampleBufferComputation):
    """GAUC: mean of per-group AUCs (groups with a single class are skipped)."""

    NAME = MetricName.GROUPED_AUC
    EXTRA = ["grouping_keys"]

    def _value(self, p, l, w, extra):
        g = extra["grouping_keys"]
        aucs = []
        for k in torch.unique(g):
            m = g == k
            if l[m].min() == l[m].max():
                continue
            aucs.append(_auc_from_samples(p[m], l[m], w[m]))
        return torch.stack(aucs).mean() if aucs else torch.tensor(0.5, dtype=torch.double)


class XAUCMetricComputation(_SampleBufferComputation):
    """Cross AUC for regression: P(pred_i > pred_j | label_i > label_j), weighted by w_i * w_j."""

    NAME = MetricName.XAUC

    def _value(self, p, l, w, extra):
        n = p.numel()
        if n < 2:
            return torch.tensor(0.0, dtype=torch.double)
        if n > 4096:
            idx = torch.randperm(n)[:4096]
            p, l, w = p[idx], l[idx], w[idx]
        ww = w.unsqueeze(0) * w.unsqueeze(1)
        dp = torch.sign(p.unsqueeze(0) - p.unsqueeze(1))
        dl = torch.sign(l.unsqueeze(0) - l.unsqueeze(1))
        match = ((dp == dl) & (dl != 0)).double()
        iu = torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1)
        return (ww * match)[iu].sum() / (ww[iu].sum() + EPS)


class NDCGMetricComputation(_SampleBufferComputation):
    """Session NDCG: samples grouped by ``session_ids``; gain = label (or 2^label - 1 with exponential_gain)."""

    NAME = MetricName.NDCG
    EXTRA = ["session_ids"]

    def __init__(self, *args: Any, exponential_gain: bool = False, k: int = -1, **kwargs: Any) -> None:
        kwargs.pop("session_key", None)
        self._exp = exponential_gain
        self._k = k
        super().__init__(*args, **kwargs)

    def _value(self, p, l, w, extra):
        s = extra["session_ids"]
        vals = []
        for sid in torch.unique(s):
            m = s == sid
            gains = (2.0 ** l[m] - 1.0) if self._exp else l[m]
            k = gains.numel() if self._k <= 0 else min(self._k, gains.numel())
            disc = 1.0 / torch.log2(torch.arange(2, k + 2, dtype=torch.double))
            dcg = (gains[torch.argsort(p[m], descending=True)][:k] * disc).sum()
            idcg = (torch.sort(gains, descending=True).values[:k] * disc).sum()
            if idcg > 0:
                vals.append(dcg / idcg)
        return torch.stack(vals).mean() if vals else torch.tensor(0.0, dtype=torch.double)


class RecallSessionMetricComputation(_SampleBufferComputation):
    """Session-level recall: per session, top-k predictions count as positive predictions."""

    NAME = MetricName.RECALL_SESSION_LEVEL
    EXTRA = ["session_ids"]

    def __init__(self, *args: Any, session_metric_def: Optional[Any] = None, top_threshold: int = 1, **kwargs: Any) -> None:
        self._top = getattr(session_metric_def, "top_threshold", None) or top_threshold
        super().__init__(*args, **kwargs)

    def _counts(self, p, l, s):
        tp = fn = fp = 0.0
        for sid in torch.unique(s):
            m = s == sid
            order = torch.argsort(p[m], descending=True)
            pred_pos = torch.zeros(int(m.sum()), dtype=torch.bool)
            pred_pos[order[: self._top]] = True
            lab = l[m] > 0
            tp += float((pred_pos & lab).sum())
            fn += float((~pred_pos & lab).sum())
            fp += float((pred_pos & ~lab).sum())
        return tp, fn, fp

    def _value(self, p, l, w, extra):
        tp, fn, fp = self._counts(p, l, extra["session_ids"])
        return torch.tensor(tp / (tp + fn) if tp + fn > 0 else 0.0, dtype=torch.double)


class PrecisionSessionMetricComputation(RecallSessionMetricComputation):
    NAME = MetricName.PRECISION_SESSION_LEVEL

    def _value(self, p, l, w, extra):
        tp, fn, fp = self._counts(p, l, extra["session_ids"])
        return torch.tensor(tp / (tp + fp) if tp + fp > 0 else 0.0, dtype=torch.double)


class SegmentedNEMetricComputation(RecMetricComputation):
    """NE per segment (``grouping_keys`` in [0, num_groups))."""

    def __init__(self, *args: Any, num_groups: int = 1, grouping_keys: str = "grouping_keys", **kwargs: Any) -> None:
        kwargs.pop("include_logloss", None)
        super().__init__(*args, **kwargs)
        self._num_groups = num_groups
        for s in ["cross_entropy_sum", "weighted_num_samples", "pos_labels", "neg_labels"]:
            self._add_state(s, torch.zeros(self._n_tasks, num_groups, dtype=torch.double), add_window_state=False, dist_reduce_fx="sum")

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        g = kwargs.get("grouping_keys")
        if g is None and "required_inputs" in kwargs:
            g = kwargs["required_inputs"].get("grouping_keys")
        g = g.reshape(1, -1).expand(self._n_tasks, -1).long()
        ce = _ce(labels, predictions, weights)
        w = weights.double()
        for name, v in (("cross_entropy_sum", ce), ("weighted_num_samples", w), ("pos_labels", w * labels.double()), ("neg_labels", w * (1 - labels.double()))):
            st = getattr(self, name)
            st.scatter_add_(1, g.to(st.device), v.to(st.device))

    def _compute(self) -> List[MetricComputationReport]:
        ne = compute_ne(self.cross_entropy_sum, self.weighted_num_samples, self.pos_labels, self.neg_labels)
        return [MetricComputationReport(MetricName.SEGMENTED_NE, MetricPrefix.LIFETIME, ne[:, gi], description=f"_{gi}") for gi in range(self._num_groups)]


class TensorWeightedAvgMetricComputation(_SumStatesComputation):
    """Weighted average of an arbitrary named tensor from ``required_inputs``."""

    STATES = ["weighted_sum", "weighted_num_samples"]

    def __init__(self, *args: Any, tensor_name: Optional[str] = None, weighted: bool = True, description: Optional[str] = None, **kwargs: Any) -> None:
        self._tensor_name = tensor_name
        self._weighted = weighted
        self._description = description
        super().__init__(*args, **kwargs)

    def _needs(self):
        return []

    def _batch_states(self, predictions, labels, weights, **kwargs):
        t = kwargs.get("required_inputs", {}).get(self._tensor_name) if self._tensor_name else predictions
        if t is None:
            raise RecMetricException(f"TensorWeightedAvg needs required input '{self._tensor_name}'")
        t = t.reshape(1, -1).double()
        w = weights.double() if self._weighted else torch.ones_like(t)
        return {"weighted_sum": (t * w).sum(-1), "weighted_num_samples": w.sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.TENSOR_WEIGHTED_AVG, prefix, get("weighted_sum") / (get("weighted_num_samples") + EPS), description=self._description)]


class UnweightedNEMetricComputation(NEMetricComputation):
    def _batch_states(self, predictions, labels, weights, **kwargs):
        return super()._batch_states(predictions, labels, torch.ones_like(weights), **kwargs)

    def _reports(self, get, prefix):
        reps = super()._reports(get, prefix)
        reps[0] = MetricComputationReport(MetricName.UNWEIGHTED_NE, prefix, reps[0].value)
        return reps


class CaliFreeNEMetricComputation(NEMetricComputation):
    """NE after rescaling predictions to perfect calibration (isolates ranking quality)."""

    STATES = NEMetricComputation.STATES + ["weighted_sum_predictions"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        st = super()._batch_states(predictions, labels, weights, **kwargs)
        st["weighted_sum_predictions"] = (weights.double() * predictions.double()).sum(-1)
        return st

    def _reports(self, get, prefix):
        ne = compute_ne(get("cross_entropy_sum"), get("weighted_num_samples"), get("pos_labels"), get("neg_labels"), self.eta)
        mean_pred = get("weighted_sum_predictions") / (get("weighted_num_samples") + EPS)
        mean_label = get("pos_labels") / (get("weighted_num_samples") + EPS)
        ent = lambda q: -(mean_label * torch.log2(q + self.eta) + (1 - mean_label) * torch.log2(1 - q + self.eta))
        cali_term = (ent(mean_pred) - ent(mean_label)) / (ent(mean_label) + EPS)
        return [MetricComputationReport(MetricName.CALI_FREE_NE, prefix, ne - cali_term)]


class ServingNEMetricComputation(NEMetricComputation):
    def _reports(self, get, prefix):
        reps = super()._reports(get, prefix)
        return [MetricComputationReport(MetricName.SERVING_NE, prefix, reps[0].value)]


class ServingCalibrationMetricComputation(CalibrationMetricComputation):
    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.SERVING_CALIBRATION, prefix, get("calibration_num") / (get("calibration_denom") + EPS))]


class OutputMetricComputation(_SumStatesComputation):
    """Mean prediction and mean label (model output monitoring)."""

    STATES = ["latest_imp", "total_latest_imp"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"latest_imp": (predictions.double() * weights.double()).sum(-1), "total_latest_imp": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.OUTPUT, prefix, get("latest_imp") / (get("total_latest_imp") + EPS))]


class HindsightTargetPRMetricComputation(RecMetricComputation):
    """Precision/recall at the threshold that reaches a target precision in hindsight (bucketed thresholds)."""

    def __init__(self, *args: Any, target_precision: float = 0.5, threshold_granularity: int = 1000, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self._target = target_precision
        self._gran = threshold_granularity
        for s in ["true_pos_sum", "false_pos_sum", "false_neg_sum"]:
            self._add_state(s, torch.zeros(self._n_tasks, threshold_granularity, dtype=torch.double), add_window_state=False, dist_reduce_fx="sum")

    def update(self, *, predictions, labels, weights, **kwargs: Any) -> None:
        th = torch.linspace(0, 1, self._gran, dtype=torch.double).view(1, -1, 1)
        pred = (predictions.double().unsqueeze(1) >= th).double()
        l, w = labels.double().unsqueeze(1), weights.double().unsqueeze(1)
        self.true_pos_sum += (w * pred * l).sum(-1).to(self.true_pos_sum.device)
        self.false_pos_sum += (w * pred * (1 - l)).sum(-1).to(self.true_pos_sum.device)
        self.false_neg_sum += (w * (1 - pred) * l).sum(-1).to(self.true_pos_sum.device)

    def _compute(self) -> List[MetricComputationReport]:
        prec = self.true_pos_sum / (self.true_pos_sum + self.false_pos_sum + EPS)
        rec = self.true_pos_sum / (self.true_pos_sum + self.false_neg_sum + EPS)
        ok = prec >= self._target
        idx = torch.where(ok.any(-1), ok.double().argmax(-1), torch.full((self._n_tasks,), self._gran - 1))
        ar = torch.arange(self._n_tasks)
        return [MetricComputationReport(MetricName.HINDSIGHT_TARGET_PR, MetricPrefix.LIFETIME, idx.double() / (self._gran - 1), description="_threshold"),
                MetricComputationReport(MetricName.HINDSIGHT_TARGET_PR, MetricPrefix.LIFETIME, prec[ar, idx], description="_precision"),
                MetricComputationReport(MetricName.HINDSIGHT_TARGET_PR, MetricPrefix.LIFETIME, rec[ar, idx], description="_recall")]


class MultiLabelPrecisionMetricComputation(PrecisionMetricComputation):
    def _reports(self, get, prefix):
        r = super()._reports(get, prefix)
        return [MetricComputationReport(MetricName.MULTI_LABEL_PRECISION, prefix, r[0].value)]


def _recalibrate(predictions: torch.Tensor, coef: float) -> torch.Tensor:
    """Undo negative down-sampling: p -> p / (p + (1 - p) / c). Parity: ne_with_recalibration.py:76-85."""
    p = predictions.double()
    return p / (p + (1.0 - p) / coef)


class RecalibratedNEMetricComputation(NEMetricComputation):
    """NE on predictions re-calibrated for the training-time negative down-sampling rate. Parity: ne_with_recalibration.py:20-116."""

    def __init__(self, *args: Any, recalibration_coefficient: float = 1.0, **kwargs: Any) -> None:
        self._recalibration_coefficient = float(recalibration_coefficient)
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return super()._batch_states(_recalibrate(predictions, self._recalibration_coefficient), labels, weights, **kwargs)

    def _reports(self, get, prefix):
        reps = super()._reports(get, prefix)
        return [MetricComputationReport(MetricName.RECALIBRATED_NE, prefix, reps[0].value)] + reps[1:]


class RecalibratedCalibrationMetricComputation(CalibrationMetricComputation):
    """Calibration of re-calibrated predictions. Parity: calibration_with_recalibration.py:25-100."""

    def __init__(self, *args: Any, recalibration_coefficient: float = 1.0, **kwargs: Any) -> None:
        self._recalibration_coefficient = float(recalibration_coefficient)
        super().__init__(*args, **kwargs)

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return super()._batch_states(_recalibrate(predictions, self._recalibration_coefficient), labels, weights, **kwargs)

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.RECALIBRATED_CALIBRATION, prefix, get("calibration_num") / (get("calibration_denom") + EPS))]


class ServingAELossMetricComputation(_SumStatesComputation):
    """Weighted mean absolute error of the served prediction (the reference only reserves the name: metrics_config.py:48)."""

    STATES = ["error_sum", "weighted_num_samples"]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {"error_sum": (weights.double() * (labels.double() - predictions.double()).abs()).sum(-1), "weighted_num_samples": weights.double().sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.SERVING_AE_LOSS, prefix, get("error_sum") / (get("weighted_num_samples") + EPS))]


class _SingleSumComputation(_SumStatesComputation):
    """One weighted sum reported as-is (data-volume monitors)."""

    NAME: MetricName

    def _needs(self):
        return []

    def _sum(self, predictions, labels, weights) -> torch.Tensor:
        raise NotImplementedError

    def _batch_states(self, predictions, labels, weights, **kwargs):
        return {self.STATES[0]: self._sum(predictions, labels, weights)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(self.NAME, prefix, get(self.STATES[0]))]


class NumPositiveSamplesMetricComputation(_SingleSumComputation):
    """sum w * label (NaN labels count 0). Parity: num_positive_samples.py:21-95."""

    STATES = ["weighted_pos_sum"]
    NAME = MetricName.NUM_POSITIVE_SAMPLES

    def _sum(self, predictions, labels, weights):
        return (weights.double() * torch.nan_to_num(labels.double(), 0.0)).sum(-1)


class SumWeightsMetricComputation(_SingleSumComputation):
    """sum w. Parity: sum_weights.py:21-95."""

    STATES = ["weighted_sum"]
    NAME = MetricName.SUM_WEIGHTS

    def _sum(self, predictions, labels, weights):
        return weights.double().sum(-1)


class NumMissingLabelsMetricComputation(_SingleSumComputation):
    """sum of weights of samples whose label is NaN. Parity: num_missing_labels.py:21-95."""

    STATES = ["missing_label_sum"]
    NAME = MetricName.NUM_MISSING_LABELS

    def _sum(self, predictions, labels, weights):
        return torch.where(torch.isnan(labels), weights.double(), torch.zeros_like(weights, dtype=torch.double)).sum(-1)


class WeightedSumPredictionsMetricComputation(_SingleSumComputation):
    """sum w * prediction (NaN predictions count 0). Parity: weighted_sum_predictions.py:21-99."""

    STATES = ["weighted_predictions_sum"]
    NAME = MetricName.WEIGHTED_SUM_PREDICTIONS

    def _needs(self):
        return ["predictions"]

    def _sum(self, predictions, labels, weights):
        return (weights.double() * torch.nan_to_num(predictions.double(), 0.0)).sum(-1)


class NEPositiveMetricComputation(_SumStatesComputation):
    """Normalized entropy of the POSITIVE samples only: ``-sum w*y*log2(p)`` over the base-rate entropy of the positives
    (reference metrics/ne_positive.py:25-70)."""

    STATES = ["cross_entropy_positive_sum", "weighted_num_samples", "pos_labels", "neg_labels"]

    def __init__(self, *args: Any, allow_missing_label_with_zero_weight: bool = False, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.eta = 1e-12

    def _batch_states(self, predictions, labels, weights, **kwargs):
        p = torch.clamp(predictions.double(), self.eta, 1 - self.eta)
        w, y = weights.double(), labels.double()
        return {"cross_entropy_positive_sum": (-w * y * torch.log2(p)).sum(-1), "weighted_num_samples": w.sum(-1),
                "pos_labels": (w * y).sum(-1), "neg_labels": (w * (1 - y)).sum(-1)}

    def _reports(self, get, prefix):
        mean_label = get("pos_labels") / (get("weighted_num_samples") + EPS)
        norm = -get("pos_labels") * torch.log2(mean_label + self.eta)
        return [MetricComputationReport(MetricName.NE_POSITIVE, prefix, get("cross_entropy_positive_sum") / (norm + EPS))]


def _make(name: str, comp: Type[RecMetricComputation], ns: MetricNamespace) -> Type[RecMetric]:
    return type(name, (RecMetric,), {"_namespace": ns, "_computation_class": comp, "__doc__": comp.__doc__})


NEMetric = _make("NEMetric", NEMetricComputation, MetricNamespace.NE)
NEPositiveMetric = _make("NEPositiveMetric", NEPositiveMetricComputation, MetricNamespace.NE_POSITIVE)
CalibrationMetric = _make("CalibrationMetric", CalibrationMetricComputation, MetricNamespace.CALIBRATION)
CTRMetric = _make("CTRMetric", CTRMetricComputation, MetricNamespace.CTR)
MSEMetric = _make("MSEMetric", MSEMetricComputation, MetricNamespace.MSE)
MAEMetric = _make("MAEMetric", MAEMetricComputation, MetricNamespace.MAE)
NMSEMetric = _make("NMSEMetric", NMSEMetricComputation, MetricNamespace.NMSE)
AccuracyMetric = _make("AccuracyMetric", AccuracyMetricComputation, MetricNamespace.ACCURACY)
PrecisionMetric = _make("PrecisionMetric", PrecisionMetricComputation, MetricNamespace.PRECISION)
RecallMetric = _make("RecallMetric", RecallMetricComputation, MetricNamespace.RECALL)
WeightedAvgMetric = _make("WeightedAvgMetric", WeightedAvgMetricComputation, MetricNamespace.WEIGHTED_AVG)
AverageMetric = _make("AverageMetric", AverageMetricComputation, MetricNamespace.AVERAGE)
ScalarMetric = _make("ScalarMetric", ScalarMetricComputation, MetricNamespace.SCALAR)
TowerQPSMetric = _make("TowerQPSMetric", TowerQPSMetricComputation, MetricNamespace.TOWER_QPS)
MulticlassRecallMetric = _make("MulticlassRecallMetric", MulticlassRecallMetricComputation, MetricNamespace.MULTICLASS_RECALL)
AUCMetric = _make("AUCMetric", AUCMetricComputation, MetricNamespace.AUC)
AUPRCMetric = _make("AUPRCMetric", AUPRCMetricComputation, MetricNamespace.AUPRC)
RAUCMetric = _make("RAUCMetric", RAUCMetricComputation, MetricNamespace.RAUC)
GAUCMetric = _make("GAUCMetric", GroupedAUCMetricComputation, MetricNamespace.GROUPED_AUC)
XAUCMetric = _make("XAUCMetric", XAUCMetricComputation, MetricNamespace.XAUC)
NDCGMetric = _make("NDCGMetric", NDCGMetricComputation, MetricNamespace.NDCG)
RecallSessionMetric = _make("RecallSessionMetric", RecallSessionMetricComputation, MetricNamespace.RECALL_SESSION_LEVEL)
PrecisionSessionMetric = _make("PrecisionSessionMetric", PrecisionSessionMetricComputation, MetricNamespace.PRECISION_SESSION_LEVEL)
SegmentedNEMetric = _make("SegmentedNEMetric", SegmentedNEMetricComputation, MetricNamespace.SEGMENTED_NE)
TensorWeightedAvgMetric = _make("TensorWeightedAvgMetric", TensorWeightedAvgMetricComputation, MetricNamespace.TENSOR_WEIGHTED_AVG)
UnweightedNEMetric = _make("UnweightedNEMetric", UnweightedNEMetricComputation, MetricNamespace.UNWEIGHTED_NE)
CaliFreeNEMetric = _make("CaliFreeNEMetric", CaliFreeNEMetricComputation, MetricNamespace.CALI_FREE_NE)
ServingNEMetric = _make("ServingNEMetric", ServingNEMetricComputation, MetricNamespace.SERVING_NE)
ServingCalibrationMetric = _make("ServingCalibrationMetric", ServingCalibrationMetricComputation, MetricNamespace.SERVING_CALIBRATION)
OutputMetric = _make("OutputMetric", OutputMetricComputation, MetricNamespace.OUTPUT)
HindsightTargetPRMetric = _make("HindsightTargetPRMetric", HindsightTargetPRMetricComputation, MetricNamespace.HINDSIGHT_TARGET_PR)
MultiLabelPrecisionMetric = _make("MultiLabelPrecisionMetric", MultiLabelPrecisionMetricComputation, MetricNamespace.MULTI_LABEL_PRECISION)
RecalibratedNEMetric = _make("RecalibratedNEMetric", RecalibratedNEMetricComputation, MetricNamespace.RECALIBRATED_NE)
RecalibratedCalibrationMetric = _make("RecalibratedCalibrationMetric", RecalibratedCalibrationMetricComputation, MetricNamespace.RECALIBRATED_CALIBRATION)
ServingAELossMetric = _make("ServingAELossMetric", ServingAELossMetricComputation, MetricNamespace.SERVING_AE_LOSS)
NumPositiveSamplesMetric = _make("NumPositiveSamplesMetric", NumPositiveSamplesMetricComputation, MetricNamespace.NUM_POSITIVE_SAMPLES)
SumWeightsMetric = _make("SumWeightsMetric", SumWeightsMetricComputation, MetricNamespace.SUM_WEIGHTS)
NumMissingLabelsMetric = _make("NumMissingLabelsMetric", NumMissingLabelsMetricComputation, MetricNamespace.NUM_MISSING_LABELS)
WeightedSumPredictionsMetric = _make("WeightedSumPredictionsMetric", WeightedSumPredictionsMetricComputation, MetricNamespace.WEIGHTED_SUM_PREDICTIONS)
