"""A RecMetric whose internal state the test controls: it records every ``update`` call, computes nothing, and lets the test read /
set / add to / append to the states of its computations - for tests of metric modules, state snapshots and asynchronous update paths
that must not depend on a real metric's arithmetic (reference ``torchrec/metrics/test_utils/mock_metrics.py``)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union
from unittest.mock import MagicMock

import torch

from ..metrics_namespace import MetricNamespaceBase
from ..rec_metric import MetricComputationReport, RecComputeMode, RecMetric, RecMetricComputation, RecTaskInfo

RecModelOutput = Union[torch.Tensor, Dict[str, torch.Tensor]]


class MockRecMetricComputation(RecMetricComputation):
    """States come from ``initial_states``: tensors (sum-like states, e.g. NE) or lists of tensors (sample buffers, e.g. AUC)."""

    def __init__(self, *args: Any, initial_states: Optional[Dict[str, Any]] = None, reduction_fn: Union[str, Callable[..., Any]] = "sum", **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        for name, value in (initial_states or {}).items():
            if isinstance(value, torch.Tensor):
                self._add_state(name, value, add_window_state=False, dist_reduce_fx=reduction_fn if isinstance(reduction_fn, str) else "sum", persistent=True)
            else:  # tensor list state: a plain attribute, tracked like a state
                setattr(self, name, list(value))
                self._state_names.append(name)
                self._reductions[name] = reduction_fn if isinstance(reduction_fn, str) else "cat"

    def update(self, *args: Any, **kwargs: Any) -> None:
        pass

    def _compute(self) -> List[MetricComputationReport]:
        return []


class MockRecMetric(RecMetric):
    _computation_class = MockRecMetricComputation
    _namespace: MetricNamespaceBase = MagicMock()

    def __init__(self, world_size: int, my_rank: int, batch_size: int, tasks: List[RecTaskInfo], compute_mode: RecComputeMode = RecComputeMode.UNFUSED_TASKS_COMPUTATION,
                 reduction_fn: Union[str, Callable[..., Any]] = "sum", initial_states: Optional[Dict[str, Any]] = None, is_tensor_list: bool = False, **kwargs: Any) -> None:
        initial_states = initial_states or create_tensor_states(["state_1", "state_2", "state_3"])
        defaults = {name: [] for name in initial_states} if is_tensor_list else initial_states
        kwargs.setdefault("window_size", max(batch_size, 100))
        super().__init__(world_size=world_size, my_rank=my_rank, batch_size=batch_size, tasks=tasks, compute_mode=compute_mode,
                         **{**kwargs, "initial_states": defaults, "reduction_fn": reduction_fn})
        if is_tensor_list:
            self.set_computation_states(initial_states)
        self.reset()

    # ---- call recording ------------------------------------------------------------------------------------------------------------------------------
    def update(self, *, predictions: RecModelOutput, labels: RecModelOutput, weights: Optional[RecModelOutput] = None, **kwargs: Any) -> None:
        self.update_called_count += 1
        self.predictions_update_calls.append(predictions)
        self.labels_update_calls.append(labels)
        self.weights_update_calls.append(weights)

    def update_called(self) -> bool:
        return self.update_called_count > 0

    def compute(self) -> Dict[str, torch.Tensor]:
        self._compute_called = True
        return {}

    def compute_called(self) -> bool:
        return self._compute_called

    def reset(self) -> None:
        self.update_called_count = 0
        self.predictions_update_calls: List[RecModelOutput] = []
        self.labels_update_calls: List[RecModelOutput] = []
        self.weights_update_calls: List[Optional[RecModelOutput]] = []
        self._compute_called = False

    # ---- state control ----------------------------------------------------------------------------------------------------------------------------------
    def get_computation_states(self) -> Dict[str, Any]:
        states: Dict[str, Any] = {}
        for comp in self._metrics_computations:
            for name in comp._reductions:
                if hasattr(comp, name):
                    states[name] = getattr(comp, name)
        return states

    def set_computation_states(self, states: Dict[str, Any]) -> None:
        for comp in self._metrics_computations:
            for name, value in states.items():
                if name in comp._reductions:
                    if isinstance(value, torch.Tensor) and name in comp._buffers:
                        comp._buffers[name] = value
                    else:
                        if name in comp._buffers:
                            del comp._buffers[name]
                        object.__setattr__(comp, name, list(value) if isinstance(value, list) else value)

    def add_to_computation_states(self, states: Dict[str, torch.Tensor]) -> None:
        for comp in self._metrics_computations:
            for name, value in states.items():
                if name in comp._reductions:
                    cur = getattr(comp, name)
                    cur += value

    def append_to_computation_states(self, states: Dict[str, torch.Tensor]) -> None:
        for comp in self._metrics_computations:
            for name, value in states.items():
                if name in comp._reductions:
                    getattr(comp, name).append(value)

    def verify_sync_disabled(self) -> bool:
        """No computation is bound to a process group (a snapshot / offloaded copy must never issue collectives)."""
        return all(comp._process_group is None for comp in self._metrics_computations)


def create_metric_states_dict(metric_prefix: str, computation_name: str, metric_states: Dict[str, Any]) -> Dict[str, Any]:
    """``{<prefix>_<computation>_<state>: value}`` - the key layout of a flattened metric-module state."""
    return {f"{metric_prefix}_{computation_name}_{name}": value for name, value in metric_states.items()}


def assert_tensor_dict_equals(actual_states: Dict[str, Any], expected_states: Dict[str, Any]) -> None:
    assert set(actual_states) == set(expected_states), f"Keys mismatch. Expected {set(expected_states)}, got {set(actual_states)}"
    for key, expected in expected_states.items():
        actual = actual_states[key]
        if isinstance(expected, torch.Tensor):
            torch.testing.assert_close(actual.to(expected.device), expected, msg=f"Mismatch for key {key}. Expected {expected}, got {actual}")
        elif isinstance(expected, list):
            assert len(actual) == len(expected), f"Length mismatch for key {key}. Expected {len(expected)}, got {len(actual)}"
            for i, (a, e) in enumerate(zip(actual, expected)):
                if isinstance(e, torch.Tensor):
                    torch.testing.assert_close(a.to(e.device), e, msg=f"Mismatch for key {key}[{i}]")
                else:
                    assert a == e, f"Mismatch for key {key}[{i}]. Expected {e}, got {a}"
        else:
            assert actual == expected, f"Mismatch for key {key}. Expected {expected}, got {actual}"


def create_tensor_states(keys: List[str], n_tasks: int = 1) -> Dict[str, Any]:
    return {key: torch.rand(n_tasks) for key in keys}


def create_tensor_list_states(keys: List[str]) -> Dict[str, Any]:
    return {key: [torch.rand(1, 2)] for key in keys}
