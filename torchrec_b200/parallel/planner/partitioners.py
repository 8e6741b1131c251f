"""Partitioners: place the shards of a proposal on devices (reference planner/partitioners.py:176-825)."""
from __future__ import annotations

import copy
import heapq
import logging
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, cast

from ..types import ShardingType
from .perf_models import NoopPerfModel
from .types import DeviceHardware, PartitionByType, Partitioner, Perf, PerfModel, PlannerError, PlannerErrorType, ShardingOption, Storage, Topology
from .utils import bytes_to_gb, reset_shard_rank

logger = logging.getLogger(__name__)


def _sort_devices_by_perf(devices: List[List[DeviceHardware]]) -> List[List[DeviceHardware]]:
    def _get_perf_sum(device_list: List[DeviceHardware]) -> float:
        return sum(d.perf.total for d in device_list)

    return sorted(devices, key=_get_perf_sum)


def _get_uniform_sharding_options(sharding_options: List[ShardingOption]) -> List[ShardingOption]:
    return [so for so in sharding_options if so.partition_by == PartitionByType.UNIFORM.value]


@dataclass
class ShardingOptionGroup:
    sharding_options: List[ShardingOption]
    storage_sum: Storage
    perf_sum: float
    param_count: int


class SortBy:
    STORAGE = "storage"
    PERF = "perf"


def _group_and_sort_non_uniform_sharding_options(sharding_options: List[ShardingOption], sort_by: str = SortBy.STORAGE,
                                                 balance_modules: bool = False) -> List[ShardingOptionGroup]:
    # options that declare a dependency are co-located: they form one group
    groups: Dict[str, ShardingOptionGroup] = {}
    for so in sharding_options:
        if so.partition_by == PartitionByType.UNIFORM.value:
            continue
        key = so.dependency or so.fqn
        if key not in groups:
            groups[key] = ShardingOptionGroup([so], copy.deepcopy(so.total_storage), so.total_perf, 1)
        else:
            g = groups[key]
            g.sharding_options.append(so)
            g.storage_sum += so.total_storage
            g.perf_sum += so.total_perf
            g.param_count += 1
    lst = list(groups.values())
    if sort_by == SortBy.PERF:
        lst.sort(key=lambda g: -g.perf_sum)
    else:
        lst.sort(key=lambda g: (-g.storage_sum.hbm, -g.storage_sum.ddr, -g.perf_sum))
    return lst


@dataclass
class OrderedDeviceHardware:
    """Heap entry of a device: least loaded first; ties go to the lower LOCAL rank (spreads tables over hosts before filling one
    host's DDR), then to the lower rank."""

    device: DeviceHardware
    local_world_size: int

    def _key(self) -> Tuple[float, int, int]:
        return (self.device.perf.total, self.device.rank % self.local_world_size, self.device.rank)

    def __lt__(self, other: "OrderedDeviceHardware") -> bool:
        return self._key() < other._key()


class GreedyPerfPartitioner(Partitioner):
    """Greedy: biggest groups first, each shard onto the currently least-loaded device (by accumulated perf)
    that still has room. HOST-partitioned options (TWRW/TWCW) choose the least-loaded host; UNIFORM options
    (RW/DP) take one shard per device; MULTI_HOST (GRID) spreads column shards over hosts."""

    def __init__(self, sort_by: str = SortBy.STORAGE, balance_modules: bool = False) -> None:
        self._sort_by = sort_by
        self._balance_modules = balance_modules

    def partition(self, proposal: List[ShardingOption], storage_constraint: Topology) -> List[ShardingOption]:
        _topology: Topology = copy.deepcopy(storage_constraint)
        minheap_devices: Optional[List] = None
        _host_level_devices = self._get_host_level_devices(_topology)
        # uniform options first: they hit every device
        uniform = _get_uniform_sharding_options(proposal)
        self._uniform_partition(uniform, _topology.devices)
        for group in _group_and_sort_non_uniform_sharding_options(proposal, self._sort_by, self._balance_modules):
            pb = group.sharding_options[0].partition_by
            if pb == PartitionByType.MULTI_HOST.value:
                for so in group.sharding_options:
                    self._multi_hosts_partition(so, _host_level_devices)
                minheap_devices = None
            elif pb == PartitionByType.HOST.value:
                self._cohost_partition(group, _host_level_devices)
                minheap_devices = None
            elif pb == PartitionByType.DEVICE.value:
                if minheap_devices is None:
                    minheap_devices = self._establish_minheap(_topology.devices, _topology.local_world_size)
                assert len(group.sharding_options) >= 1
                if len(group.sharding_options) == 1:
                    self._device_partition(group.sharding_options[0], minheap_devices)
                else:
                    self._cogroup_device_partition(group, minheap_devices)
            else:
                raise RuntimeError(f"Unexpected sharding option group {group}")
        self._topology = _topology
        return proposal

    # ---- helpers --------------------------------------------------------------------------------------
    @staticmethod
    def _establish_minheap(devices: List[DeviceHardware], local_world_size: int) -> List:
        heap = [[d.perf.total, d.rank, d] for d in devices]
        heapq.heapify(heap)
        return heap

    @staticmethod
    def _fits(shard_storage: Storage, device: DeviceHardware) -> bool:
        return shard_storage.hbm <= device.storage.hbm and shard_storage.ddr <= device.storage.ddr

    @classmethod
    def _device_partition(cls, sharding_option: ShardingOption, minheap_devices: List, bulk_heapify_threshold: float = 0.25) -> None:
        # largest shards first
        order = sorted(range(len(sharding_option.shards)), key=lambda i: -cast(Storage, sharding_option.shards[i].storage).hbm)
        for i in order:
            shard = sharding_option.shards[i]
            tmp = []
            placed = False
            while minheap_devices:
                entry = heapq.heappop(minheap_devices)
                device = entry[2]
                if cls._fits(cast(Storage, shard.storage), device):
                    shard.rank = device.rank
                    device.storage -= cast(Storage, shard.storage)
                    device.perf += cast(Perf, shard.perf)
                    entry[0] = device.perf.total
                    heapq.heappush(minheap_devices, entry)
                    placed = True
                    break
                tmp.append(entry)
            for e in tmp:
                heapq.heappush(minheap_devices, e)
            if not placed:
                raise PlannerError(error_type=PlannerErrorType.PARTITION,
                                   message=f"Device partition failed. Couldn't find a rank for shard {shard} of table {sharding_option.name}, "
                                           f"largest device storage: {max((e[2].storage for e in minheap_devices), key=lambda s: s.hbm) if minheap_devices else None}")

    @classmethod
    def _cogroup_device_partition(cls, group: ShardingOptionGroup, minheap_devices: List) -> None:
        # all tables of the group must land on the same device(s), shard k of every table on the same rank
        n = max(so.num_shards for so in group.sharding_options)
        for k in range(n):
            shards = [so.shards[k] for so in group.sharding_options if k < so.num_shards]
            need = Storage(0, 0)
            perf = Perf()
            for s in shards:
                need += cast(Storage, s.storage)
                perf += cast(Perf, s.perf)
            tmp, placed = [], False
            while minheap_devices:
                entry = heapq.heappop(minheap_devices)
                device = entry[2]
                if cls._fits(need, device):
                    for s in shards:
                        s.rank = device.rank
                    device.storage -= need
                    device.perf += perf
                    entry[0] = device.perf.total
                    heapq.heappush(minheap_devices, entry)
                    placed = True
                    break
                tmp.append(entry)
            for e in tmp:
                heapq.heappush(minheap_devices, e)
            if not placed:
                raise PlannerError(error_type=PlannerErrorType.PARTITION, message=f"can't place co-located group {[so.name for so in group.sharding_options]}")

    @classmethod
    def _cohost_partition(cls, group: ShardingOptionGroup, _host_level_devices: List[List[DeviceHardware]]) -> None:
        sorted_hosts = _sort_devices_by_perf(_host_level_devices)
        for devices in sorted_hosts:
            host_devices = copy.deepcopy(devices)
            host_storage = Storage(sum(d.storage.hbm for d in host_devices), sum(d.storage.ddr for d in host_devices))
            if not group.storage_sum.fits_in(host_storage):
                continue
            success = True
            for so in group.sharding_options:
                try:
                    if so.sharding_type == ShardingType.TABLE_ROW_WISE.value:
                        cls._uniform_partition([so], host_devices)
                    elif so.sharding_type == ShardingType.TABLE_COLUMN_WISE.value:
                        cls._device_partition(so, cls._establish_minheap(host_devices, len(host_devices)))
                    else:
                        raise RuntimeError(f"unexpected cohost sharding type: {so.sharding_type}")
                except PlannerError:
                    success = False
                    break
            if success:
                for d, hd in zip(devices, host_devices):
                    d.storage = hd.storage
                    d.perf = hd.perf
                return
            for so in group.sharding_options:
                for shard in so.shards:
                    shard.rank = None
        raise PlannerError(error_type=PlannerErrorType.PARTITION, message=f"can't find a host for sharding option group {[so.name for so in group.sharding_options]}")

    @classmethod
    def _multi_hosts_partition(cls, sharding_option: ShardingOption, _host_level_devices: List[List[DeviceHardware]]) -> None:
        """GRID: consecutive runs of ``local_world_size`` shards are the row shards of one column shard -> one host each."""
        local = len(_host_level_devices[0])
        n_col = sharding_option.num_shards // local
        hosts = _sort_devices_by_perf(_host_level_devices)
        if n_col > len(hosts):
            raise PlannerError(error_type=PlannerErrorType.PARTITION, message=f"grid shard of {sharding_option.name} needs {n_col} hosts")
        for ci in range(n_col):
            devices = hosts[ci]
            for ri in range(local):
                shard = sharding_option.shards[ci * local + ri]
                d = devices[ri]
                if not cls._fits(cast(Storage, shard.storage), d):
                    raise PlannerError(error_type=PlannerErrorType.PARTITION, message=f"grid shard {shard} does not fit on rank {d.rank}")
                shard.rank = d.rank
                d.storage -= cast(Storage, shard.storage)
                d.perf += cast(Perf, shard.perf)

    @staticmethod
    def _get_host_level_devices(_topology: Topology) -> List[List[DeviceHardware]]:
        num_hosts: int = _topology.world_size // _topology.local_world_size
        return [_topology.devices[i * _topology.local_world_size : (i + 1) * _topology.local_world_size] for i in range(num_hosts)]

    @staticmethod
    def _uniform_partition(sharding_options: List[ShardingOption], devices: List[DeviceHardware]) -> None:
        for so in sharding_options:
            if so.num_shards != len(devices):
                raise PlannerError(error_type=PlannerErrorType.PARTITION,
                                   message=f"For a uniform partition, the number of shards ({so.num_shards}) must equal the number of devices ({len(devices)})")
            for i in range(len(devices)):
                storage_needed = cast(Storage, so.shards[i].storage)
                if not storage_needed.fits_in(devices[i].storage):
                    raise PlannerError(error_type=PlannerErrorType.PARTITION,
                                       message=f"Shard of size {storage_needed} bytes does not fit on any rank. Device memory cap: {devices[i].storage}.")
                so.shards[i].rank = devices[i].rank
                devices[i].storage -= storage_needed
                devices[i].perf += cast(Perf, so.shards[i].perf)


class MemoryBalancedPartitioner(Partitioner):
    """Repeatedly tightens the per-device HBM budget around GreedyPerfPartitioner to find a plan with lower peak
    memory at bounded perf loss (reference partitioners.py:694-825)."""

    def __init__(self, max_search_count: int = 10, tolerance: float = 0.02, balance_modules: bool = False) -> None:
        self._max_search_count = max_search_count
        self._tolerance = tolerance
        self._balance_modules = balance_modules

    def partition(self, proposal: List[ShardingOption], storage_constraint: Topology) -> List[ShardingOption]:
        _perf_model: PerfModel = NoopPerfModel(storage_constraint)
        _partitioner = GreedyPerfPartitioner(sort_by=SortBy.PERF, balance_modules=self._balance_modules)
        default_plan = copy.deepcopy(_partitioner.partition(proposal, storage_constraint))
        original_plan_perf = _perf_model.rate(default_plan)
        max_hbm_per_device: int = self._max_hbm_used(default_plan, storage_constraint)
        hbm_requirement: int = sum(so.total_storage.hbm for so in proposal)
        min_hbm_per_device: int = int(hbm_requirement / max(storage_constraint.world_size, 1))
        search_count = 0
        best = default_plan
        while search_count < self._max_search_count and max_hbm_per_device > min_hbm_per_device + 1:
            search_count += 1
            reset_shard_rank(proposal)
            mid = (max_hbm_per_device + min_hbm_per_device) // 2
            set_hbm_per_device(storage_constraint, mid)
            try:
                new_plan = _partitioner.partition(proposal, storage_constraint)
                new_perf = _perf_model.rate(new_plan)
                if new_perf > original_plan_perf * (1 + self._tolerance):
                    min_hbm_per_device = mid  # too much perf loss
                else:
                    best = copy.deepcopy(new_plan)
                    max_hbm_per_device = mid
            except PlannerError:
                min_hbm_per_device = mid
        for so_dst, so_src in zip(proposal, best):
            for s_dst, s_src in zip(so_dst.shards, so_src.shards):
                s_dst.rank = s_src.rank
        return proposal

    @staticmethod
    def _max_hbm_used(plan: List[ShardingOption], topology: Topology) -> int:
        used = [0] * topology.world_size
        for so in plan:
            for shard in so.shards:
                if shard.rank is not None:
                    used[shard.rank] += cast(Storage, shard.storage).hbm
        return max(used) if used else 0


def set_hbm_per_device(storage_constraint: Topology, hbm_per_device: int) -> None:
    for device in storage_constraint.devices:
        device.storage.hbm = hbm_per_device
