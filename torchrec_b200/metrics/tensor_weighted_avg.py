"""Weighted average of a named tensor of the model output.

Reference module: ``torchrec/metrics/tensor_weighted_avg.py``. The computation (states, update, reports) and the ``RecMetric`` class of this metric, on the shared bases of ``_bases.py``, plus the stateless
``compute_*`` / ``get_*_states`` helpers of the reference module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ._bases import EPS, _SumStatesComputation, _make  # noqa: F401
from .metrics_namespace import MetricName, MetricNamespace, MetricPrefix  # noqa: F401
from .rec_metric import MetricComputationReport, RecMetric, RecMetricComputation, RecMetricException  # noqa: F401
import time
from typing import Any, Type  # noqa: F401


class TensorWeightedAvgMetricComputation(_SumStatesComputation):
    """Weighted average of a tensor handed over in ``required_inputs`` (the sibling of ``WeightedAvg``, which averages the predictions).
    Every task names its tensor (``RecTaskInfo.tensor_name``) and says whether the example weights apply (``RecTaskInfo.weighted``);
    with fused tasks the per-task tensors arrive stacked as ``required_inputs['target_tensor']``. Reported as ``weighted_avg``."""

    STATES = ["weighted_sum", "weighted_num_samples"]

    def __init__(self, *args: Any, tasks: Optional[List[Any]] = None, tensor_name: Optional[str] = None, weighted: bool = True, description: Optional[str] = None,
                 **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if tasks is None:  # direct construction: one tensor name for all tasks
            from .rec_metric import RecTaskInfo

            tasks = [RecTaskInfo(name=f"task_{i}", tensor_name=tensor_name, weighted=weighted) for i in range(self._n_tasks)]
        self.tasks = list(tasks)
        for task in self.tasks:
            if task.tensor_name is None:
                raise RecMetricException("TensorWeightedAvgMetricComputation expects all tasks to have tensor_name, but got None.")
        self._description = description
        self.weighted_mask = torch.tensor([bool(t.weighted) for t in self.tasks]).unsqueeze(-1)

    def _needs(self):
        return []

    def _batch_states(self, predictions, labels, weights, **kwargs):
        req = kwargs.get("required_inputs")
        if req is None:
            raise RecMetricException("TensorWeightedAvgMetricComputation expects 'required_inputs' to exist.")
        if len(self.tasks) > 1 and "target_tensor" in req:
            target = req["target_tensor"]
        elif len(self.tasks) > 1:
            missing = [t.tensor_name for t in self.tasks if t.tensor_name not in req]
            if missing:
                raise RecMetricException(f"TensorWeightedAvgMetricComputation expects required_inputs to contain target tensors {missing}")
            target = torch.stack([req[t.tensor_name].reshape(-1) for t in self.tasks])
        else:
            name = self.tasks[0].tensor_name
            if name not in req:
                raise RecMetricException(f"TensorWeightedAvgMetricComputation expects required_inputs to contain target tensor {name}")
            target = req[name]
        target = target.reshape(len(self.tasks), -1).double()
        w = weights.double().reshape(len(self.tasks), -1)
        mask = self.weighted_mask.to(target.device)
        return {"weighted_sum": torch.where(mask, target * w, target).sum(-1), "weighted_num_samples": torch.where(mask, w, torch.ones_like(w)).sum(-1)}

    def _reports(self, get, prefix):
        return [MetricComputationReport(MetricName.WEIGHTED_AVG, prefix, get_mean(get("weighted_sum"), get("weighted_num_samples")), description=self._description)]


class TensorWeightedAvgMetric(RecMetric):
    _namespace: MetricNamespace = MetricNamespace.WEIGHTED_AVG
    _computation_class = TensorWeightedAvgMetricComputation

    def _get_task_kwargs(self, task_config) -> Dict[str, Any]:
        return {"tasks": [task_config] if not isinstance(task_config, (list, tuple)) else list(task_config)}

    def _get_task_required_inputs(self, task_config) -> set:
        """The target tensors of the tasks; one tensor must not be registered both weighted and unweighted."""
        seen: Dict[str, bool] = {}
        for task in ([task_config] if not isinstance(task_config, (list, tuple)) else task_config):
            if task.tensor_name is None:
                continue
            if task.tensor_name in seen and seen[task.tensor_name] is not task.weighted:
                raise RecMetricException(f"This target tensor was already registered as weighted={seen[task.tensor_name]}. "
                                         f"This target tensor cannot be re-registered with weighted={task.weighted}")
            seen[str(task.tensor_name)] = task.weighted
        return set(seen)


def get_mean(value_sum: torch.Tensor, num_samples: torch.Tensor) -> torch.Tensor:
    return value_sum / num_samples
