"""Session-level precision of the top-ranked items.

Reference module: ``torchrec/metrics/precision_session.py``: the session ranking of ``recall_session.py`` with the states ``num_true_pos`` /
``num_false_pos``; NaN until a positive prediction has been seen."""
from __future__ import annotations

from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport
from .recall_session import (NUM_FALSE_POS, NUM_TRUE_POS, _calc_num_false_pos, _calc_num_true_pos, _ratio_or_nan, _SessionRecMetric, _SessionTopKComputation,  # noqa: F401
                             ranking_within_session)


class PrecisionSessionMetricComputation(_SessionTopKComputation):
    STATES = [NUM_TRUE_POS, NUM_FALSE_POS]

    def _batch_states(self, predictions, labels, weights, **kwargs):
        labels, pred_pos, weights = self._labelled(predictions, labels, weights, kwargs)
        return {NUM_TRUE_POS: _calc_num_true_pos(labels, pred_pos, weights), NUM_FALSE_POS: _calc_num_false_pos(labels, pred_pos, weights)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.PRECISION_SESSION_LEVEL, prefix, _ratio_or_nan(get(NUM_TRUE_POS), get(NUM_FALSE_POS)))]


class PrecisionSessionMetric(_SessionRecMetric):
    _namespace: MetricNamespace = MetricNamespace.PRECISION_SESSION_LEVEL
    _computation_class = PrecisionSessionMetricComputation
