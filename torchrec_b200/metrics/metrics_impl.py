"""All metric implementations in one namespace (``from torchrec_b200.metrics import metrics_impl as M``): every name is defined in its own
module (``ne.py``, ``auc.py``, ``calibration.py``, ...) and re-exported here; the shared bases are in ``_bases.py``."""
from ._bases import *  # noqa: F401,F403
from ._bases import EPS, _make, _SampleBufferComputation, _SingleSumComputation, _SumStatesComputation, _zeros  # noqa: F401
from .accuracy import AccuracyMetric, AccuracyMetricComputation  # noqa: F401
from .auc import AUCMetric, AUCMetricComputation, _auc_from_samples  # noqa: F401
from .auprc import AUPRCMetric, AUPRCMetricComputation, _auprc_from_samples  # noqa: F401
from .average import AverageMetric, AverageMetricComputation  # noqa: F401
from .cali_free_ne import CaliFreeNEMetric, CaliFreeNEMetricComputation  # noqa: F401
from .calibration import CalibrationMetric, CalibrationMetricComputation  # noqa: F401
from .calibration_with_recalibration import RecalibratedCalibrationMetric, RecalibratedCalibrationMetricComputation  # noqa: F401
from .ctr import CTRMetric, CTRMetricComputation  # noqa: F401
from .gauc import GAUCMetric, GAUCMetricComputation  # noqa: F401
from .hindsight_target_pr import HindsightTargetPRMetric, HindsightTargetPRMetricComputation  # noqa: F401
from .mae import MAEMetric, MAEMetricComputation  # noqa: F401
from .mse import MSEMetric, MSEMetricComputation  # noqa: F401
from .multi_label_precision import MultiLabelPrecisionMetric, MultiLabelPrecisionMetricComputation  # noqa: F401
from .multiclass_recall import MulticlassRecallMetric, MulticlassRecallMetricComputation  # noqa: F401
from .ndcg import NDCGMetric, NDCGMetricComputation  # noqa: F401
from .ne import _ce, NEMetric, NEMetricComputation, compute_ne  # noqa: F401
from .ne_positive import NEPositiveMetric, NEPositiveMetricComputation  # noqa: F401
from .ne_with_recalibration import RecalibratedNEMetric, RecalibratedNEMetricComputation, _recalibrate  # noqa: F401
from .nmse import NMSEMetric, NMSEMetricComputation  # noqa: F401
from .num_missing_labels import NumMissingLabelsMetric, NumMissingLabelsMetricComputation  # noqa: F401
from .num_positive_samples import NumPositiveSamplesMetric, NumPositiveSamplesMetricComputation  # noqa: F401
from .output import OutputMetric, OutputMetricComputation  # noqa: F401
from .precision import PrecisionMetric, PrecisionMetricComputation  # noqa: F401
from .precision_session import PrecisionSessionMetric, PrecisionSessionMetricComputation  # noqa: F401
from .rauc import RAUCMetric, RAUCMetricComputation  # noqa: F401
from .recall import RecallMetric, RecallMetricComputation  # noqa: F401
from .recall_session import RecallSessionMetric, RecallSessionMetricComputation  # noqa: F401
from .scalar import ScalarMetric, ScalarMetricComputation  # noqa: F401
from .segmented_ne import SegmentedNEMetric, SegmentedNEMetricComputation  # noqa: F401
from .serving_ae_loss import ServingAELossMetric, ServingAELossMetricComputation  # noqa: F401
from .serving_calibration import ServingCalibrationMetric, ServingCalibrationMetricComputation  # noqa: F401
from .serving_ne import ServingNEMetric, ServingNEMetricComputation  # noqa: F401
from .sum_weights import SumWeightsMetric, SumWeightsMetricComputation  # noqa: F401
from .tensor_weighted_avg import TensorWeightedAvgMetric, TensorWeightedAvgMetricComputation  # noqa: F401
from .tower_qps import TowerQPSMetric, TowerQPSMetricComputation  # noqa: F401
from .unweighted_ne import UnweightedNEMetric, UnweightedNEMetricComputation  # noqa: F401
from .weighted_avg import WeightedAvgMetric, WeightedAvgMetricComputation  # noqa: F401
from .weighted_sum_predictions import WeightedSumPredictionsMetric, WeightedSumPredictionsMetricComputation  # noqa: F401
from .xauc import XAUCMetric, XAUCMetricComputation  # noqa: F401
