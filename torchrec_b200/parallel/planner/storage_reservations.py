"""HBM reservations for everything that is not an embedding shard (reference planner/storage_reservations.py)."""
from __future__ import annotations

import copy
import logging
import math
from typing import Dict, List, Optional, Set, Tuple

import torch
from torch import nn

from ..types import ModuleSharder
from .constants import BIGINT_DTYPE, POOLING_FACTOR
from .types import ParameterConstraints, PlannerError, PlannerErrorType, Storage, StorageReservation, Topology
from .utils import sharder_name

logger = logging.getLogger(__name__)


def _get_module_size(module: nn.Module, multiplier: float) -> int:
    parameters_size = sum(multiplier * parameter.element_size() * parameter.nelement() for parameter in module.parameters())
    buffers_size = sum(buffer.element_size() * buffer.nelement() for buffer in module.buffers())
    return round(parameters_size + buffers_size)


def _get_dense_tensor_size(module: nn.Module, shardable_modules: Set[nn.Module], multiplier: float = 6.0) -> int:
    dense_tensor_size = _get_module_size(module, multiplier) - sum(_get_module_size(m, multiplier) for m in shardable_modules)
    return max(dense_tensor_size, 0)


def _reserve_storage_percentage(topology: Topology, percent: float) -> None:
    for device in topology.devices:
        device.storage.hbm = int((1 - percent) * device.storage.hbm)


def _get_batch_inputs_and_shardable_parameters(module: nn.Module, sharders: List[ModuleSharder[nn.Module]], batch_size: int,
                                               constraints: Optional[Dict[str, ParameterConstraints]] = None) -> Tuple[List[float], Set[nn.Module]]:
    sharder_map: Dict[str, ModuleSharder[nn.Module]] = {sharder_name(sharder.module_type): sharder for sharder in sharders}
    input_lengths: List[float] = []
    batch_sizes: List[int] = []
    shardable_modules: Set[nn.Module] = set()

    def populate(m: nn.Module) -> None:
        sharder = sharder_map.get(sharder_name(type(m)), None)
        if not sharder:
            for c in m.children():
                populate(c)
            return
        names = sharder.shardable_parameters(m).keys()
        shardable_modules.add(m)
        for name in names:
            pc = constraints.get(name) if constraints else None
            lens = list(pc.pooling_factors) if pc and pc.pooling_factors else [POOLING_FACTOR]
            input_lengths.extend(lens)
            batch_sizes.extend(pc.batch_sizes if pc and pc.batch_sizes else [batch_size] * len(lens))

    populate(module)
    batch_inputs = [l * b for l, b in zip(input_lengths, batch_sizes)]
    return batch_inputs, shardable_modules


class FixedPercentageStorageReservation(StorageReservation):
    def __init__(self, percentage: float) -> None:
        assert 0 <= percentage <= 1
        self._percentage: float = percentage
        self._last_reserved_topology: Optional[Topology] = None

    def reserve(self, topology, batch_size, module, sharders, constraints=None) -> Topology:
        reserved_topology = copy.deepcopy(topology)
        _reserve_storage_percentage(reserved_topology, self._percentage)
        self._last_reserved_topology = copy.deepcopy(reserved_topology)
        return reserved_topology

    @property
    def last_reserved_topology(self) -> Optional[Topology]:
        return self._last_reserved_topology


class HeuristicalStorageReservation(StorageReservation):
    """percentage + dense params x multiplier (params, grads, optimizer state, DDP buckets) + KJT input buffers."""

    def __init__(self, percentage: float, parameter_multiplier: float = 6.0, dense_tensor_estimate: Optional[int] = None) -> None:
        assert 0 <= percentage <= 1
        self._percentage = percentage
        self._parameter_multiplier = parameter_multiplier
        self._dense_tensor_estimate = dense_tensor_estimate
        self._dense_storage: Optional[Storage] = None
        self._kjt_storage: Optional[Storage] = None
        self._last_reserved_topology: Optional[Topology] = None

    def reserve(self, topology, batch_size, module, sharders, constraints=None) -> Topology:
        reserved_topology = copy.deepcopy(topology)
        batch_inputs, shardable_modules = _get_batch_inputs_and_shardable_parameters(module, sharders, batch_size, constraints)
        _reserve_storage_percentage(reserved_topology, self._percentage)
        dense_bytes = self._dense_tensor_estimate if self._dense_tensor_estimate is not None else _get_dense_tensor_size(
            module, shardable_modules, self._parameter_multiplier)
        kjt_bytes = math.ceil(sum(batch_inputs) * BIGINT_DTYPE) * 20  # in-flight pipelined batches + a2a staging
        on_cuda = reserved_topology.compute_device == "cuda"
        self._dense_storage = Storage(hbm=dense_bytes if on_cuda else 0, ddr=0 if on_cuda else dense_bytes)
        self._kjt_storage = Storage(hbm=kjt_bytes if on_cuda else 0, ddr=0 if on_cuda else kjt_bytes)
        for device in reserved_topology.devices:
            device.storage -= self._dense_storage
            device.storage -= self._kjt_storage
            if device.storage.hbm < 0 or device.storage.ddr < 0:
                raise PlannerError(error_type=PlannerErrorType.INSUFFICIENT_STORAGE,
                                   message=f"The reserved storage (dense {self._dense_storage}, kjt {self._kjt_storage}) exceeds the device capacity")
        self._last_reserved_topology = copy.deepcopy(reserved_topology)
        return reserved_topology

    @property
    def last_reserved_topology(self) -> Optional[Topology]:
        return self._last_reserved_topology


class InferenceStorageReservation(StorageReservation):
    def __init__(self, percentage: float, dense_tensor_estimate: Optional[int] = None) -> None:
        self._inner = HeuristicalStorageReservation(percentage, parameter_multiplier=1.0, dense_tensor_estimate=dense_tensor_estimate)

    def reserve(self, topology, batch_size, module, sharders, constraints=None) -> Topology:
        return self._inner.reserve(topology, batch_size, module, sharders, constraints)

    @property
    def last_reserved_topology(self) -> Optional[Topology]:
        return self._inner.last_reserved_topology
