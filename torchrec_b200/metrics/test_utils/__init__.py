"""Test doubles for code that drives RecMetrics (metric modules, snapshots, offloaded updates)."""
from .mock_metrics import (MockRecMetric, MockRecMetricComputation, assert_tensor_dict_equals, create_metric_states_dict, create_tensor_list_states,  # noqa: F401
                           create_tensor_states)
