"""EmbeddingShardingPlanner: enumerate -> propose -> partition -> rate loop (reference planner/planners.py:148-1120)."""
from __future__ import annotations

import copy
import hashlib
import logging
import time
from functools import reduce
from typing import Callable, Dict, List, Optional, Tuple, Union, cast

import torch
import torch.distributed as dist
from torch import nn

from ..collective_utils import invoke_on_rank_and_broadcast_result
from ..comm import get_local_size
from ..sharding_plan import get_default_sharders, placement
from ..types import EmbeddingModuleShardingPlan, EnumerableShardingSpec, ModuleSharder, ParameterSharding, ShardingPlan, ShardingPlanner, ShardingType, ShardMetadata
from .constants import BATCH_SIZE, MAX_SIZE
from .enumerators import EmbeddingEnumerator
from .partitioners import GreedyPerfPartitioner, MemoryBalancedPartitioner
from .perf_models import NoopPerfModel
from .proposers import DynamicProgrammingProposer, GreedyProposer, GridSearchProposer, UniformProposer
from .stats import EmbeddingStats
from .storage_reservations import HeuristicalStorageReservation
from .types import (
    Enumerator,
    ParameterConstraints,
    Partitioner,
    PerfModel,
    PlannerError,
    PlannerErrorType,
    Proposer,
    ShardingOption,
    Stats,
    Storage,
    StorageReservation,
    Topology,
)
from .utils import bytes_to_gb, reset_shard_rank, storage_repr_in_gb

logger = logging.getLogger(__name__)


def to_sharding_plan(sharding_options: List[ShardingOption], topology: Topology) -> ShardingPlan:
    """Partitioned sharding options -> ShardingPlan (module path -> table -> ParameterSharding)."""
    compute_device = topology.compute_device
    local_size = topology.local_world_size
    plan: Dict[str, EmbeddingModuleShardingPlan] = {}
    for so in sharding_options:
        shards = so.shards
        sharding_type = so.sharding_type
        module_plan = plan.get(so.path, EmbeddingModuleShardingPlan())
        module_plan[so.name] = ParameterSharding(
            sharding_spec=None if sharding_type == ShardingType.DATA_PARALLEL.value else EnumerableShardingSpec([
                ShardMetadata(shard_sizes=shard.size, shard_offsets=shard.offset, placement=placement(compute_device, cast(int, shard.rank), local_size))
                for shard in shards
            ]),
            sharding_type=sharding_type,
            compute_kernel=so.compute_kernel,
            ranks=[cast(int, shard.rank) for shard in shards],
            cache_params=so.cache_params,
            enforce_hbm=so.enforce_hbm,
            stochastic_rounding=so.stochastic_rounding,
            bounds_check_mode=so.bounds_check_mode,
            output_dtype=so.output_dtype,
            key_value_params=so.key_value_params,
        )
        plan[so.path] = module_plan
    return ShardingPlan(plan)


def _merge_plans(best_plans: List[ShardingPlan]) -> ShardingPlan:
    if len(best_plans) == 1:
        return best_plans[0]
    merged = ShardingPlan({})
    for p in best_plans:
        merged.plan.update(p.plan)
    return merged


def validate_rank_assignment(sharding_plan: ShardingPlan, topology: Topology) -> None:
    """Every shard of the plan sits on a rank in [0, world size)."""
    for module_plan in sharding_plan.plan.values():
        for param_plan in module_plan.values():
            if param_plan.sharding_spec is None:
                continue
            for shard in param_plan.sharding_spec.shards:
                rank = shard.placement.rank()
                if rank is None or rank < 0:
                    raise PlannerError(error_type=PlannerErrorType.INVALID_RANK_ASSIGNMENT, message=f"Rank is not assigned for shard {shard}")
                if rank >= topology.world_size:
                    raise PlannerError(error_type=PlannerErrorType.INVALID_RANK_ASSIGNMENT,
                                       message=f"Shard {shard} has rank {rank} which is not below the world size {topology.world_size}.")


def validate_compute_kernels(best_plan: List[ShardingOption]) -> None:
    """The kernels of the chosen options exist and fit the rest of the option: the dense kernel only for data-parallel tables, a cache
    load factor strictly between 0 and 1 for the caching kernels, no negative storage."""
    from ..embedding_types import EmbeddingComputeKernel
    from ..types import ShardingType

    valid = {k.value for k in EmbeddingComputeKernel}
    caching = {k.value for k in EmbeddingComputeKernel if "caching" in k.value}
    violations: List[str] = []
    for so in best_plan:
        fqn, kernel = so.fqn, so.compute_kernel
        if kernel not in valid:
            violations.append(f"{fqn}: unknown compute kernel '{kernel}'")
            continue
        if kernel == EmbeddingComputeKernel.DENSE.value and so.sharding_type != ShardingType.DATA_PARALLEL.value:
            violations.append(f"{fqn}: DENSE kernel requires DATA_PARALLEL sharding, got '{so.sharding_type}'")
        if kernel in caching:
            clf = so.cache_load_factor
            if clf is not None and (clf <= 0 or clf >= 1):
                violations.append(f"{fqn}: {kernel} requires cache_load_factor strictly between 0 and 1, got {clf}")
        storage = so.total_storage
        if storage.hbm < 0:
            violations.append(f"{fqn}: {kernel} has negative HBM storage ({storage.hbm})")
        if storage.ddr < 0:
            violations.append(f"{fqn}: {kernel} has negative DDR storage ({storage.ddr})")
    if violations:
        raise PlannerError(error_type=PlannerErrorType.INVALID_COMPUTE_KERNEL,
                           message=f"Compute kernel validation failed with {len(violations)} violation(s):\n" + "\n".join(f"  - {v}" for v in violations))


def validate_modules_inclusion_in_sharding_plan(sharding_plan: ShardingPlan, module: nn.Module, sharders: List[ModuleSharder[nn.Module]],
                                                constraints: Optional[Dict[str, ParameterConstraints]] = None, device_group: Optional[str] = None) -> None:
    """Every module that has a sharder AND parameters that sharder would shard appears in the plan (with ``device_group``: only the
    modules whose tables are constrained to that group)."""
    from .utils import sharder_name

    sharder_map = {sharder_name(s.module_type): s for s in sharders}
    expected = set()
    queue: List[Tuple[str, nn.Module]] = [("", module)]
    while queue:
        path, child = queue.pop(0)
        sharder = sharder_map.get(sharder_name(type(child)))
        params = sharder.shardable_parameters(child) if sharder else None
        if params:
            in_group = device_group is None or constraints is None or any(
                getattr(constraints.get(n), "device_group", None) in (None, device_group) for n in params)
            if in_group:
                expected.add(path)
            continue
        for n, m in child.named_children():
            queue.append((f"{path}.{n}" if path else n, m))
    missing = sorted(expected - set(sharding_plan.plan.keys()))
    if missing:
        group = f" for device group '{device_group}'" if device_group else ""
        raise PlannerError(error_type=PlannerErrorType.MISSING_MODULE_IN_PLAN, message=f"The following shardable modules are not present in the sharding plan{group}: {missing}.")


def extract_plan(search_space: List[ShardingOption], loaded_sharding_options: Dict[int, ShardingOption]) -> List[ShardingOption]:
    """The options of the enumerated search space a stored plan chose, carrying the stored shards (placement included); every stored
    option must be found, and no two enumerated options may share a storage hash."""
    chosen: List[ShardingOption] = []
    seen = set()
    for so in search_space:
        h = so.storage_hash()
        if h in seen:
            raise PlannerError(error_type=PlannerErrorType.PLAN_LOADING_FAILED, message=f"Found a duplicate storage hash {h} for FQNs {[x.fqn for x in search_space]}\n")
        seen.add(h)
        loaded = loaded_sharding_options.get(h)
        if loaded is not None:
            picked = copy.copy(so)
            picked.shards = loaded.shards
            chosen.append(picked)
    if len(chosen) != len(loaded_sharding_options):
        raise PlannerError(error_type=PlannerErrorType.PLAN_LOADING_FAILED,
                           message=f"Loaded sharding options from Storage, but not all search space is covered. Merged search space len {len(chosen)} != "
                                   f"loaded Sharding options len {len(loaded_sharding_options)}\n")
    return chosen


class EmbeddingPlannerBase(ShardingPlanner):
    def __init__(self, topology: Optional[Topology] = None, batch_size: Optional[int] = None, enumerator: Optional[Enumerator] = None,
                 storage_reservation: Optional[StorageReservation] = None, stats: Optional[Union[Stats, List[Stats]]] = None,
                 constraints: Optional[Dict[str, ParameterConstraints]] = None, debug: bool = True,
                 callbacks: Optional[List[Callable[[List[ShardingOption]], List[ShardingOption]]]] = None, timeout_seconds: Optional[int] = None) -> None:
        if topology is None:
            topology = Topology(local_world_size=get_local_size(), world_size=dist.get_world_size() if dist.is_initialized() else 1,
                                compute_device="cuda" if torch.cuda.is_available() else "cpu")
        self._topology: Topology = topology
        self._batch_size: int = batch_size if batch_size else BATCH_SIZE
        self._constraints = constraints
        self._enumerator: Enumerator = enumerator if enumerator else EmbeddingEnumerator(topology=topology, batch_size=self._batch_size, constraints=constraints)
        self._storage_reservation: StorageReservation = storage_reservation if storage_reservation else HeuristicalStorageReservation(percentage=0.15)
        if stats is not None:
            self._stats: List[Stats] = [stats] if not isinstance(stats, list) else stats
        else:
            self._stats = [EmbeddingStats()]
        self._debug = debug
        self._callbacks = callbacks if callbacks is not None else []
        self._timeout_seconds = timeout_seconds
        if timeout_seconds is not None:
            assert timeout_seconds > 0, "Timeout must be positive"

    def collective_plan(self, module: nn.Module, sharders: Optional[List[ModuleSharder[nn.Module]]] = None, pg: Optional[dist.ProcessGroup] = None) -> ShardingPlan:
        """Rank 0 plans, everybody receives the same plan (object broadcast)."""
        if pg is None:
            assert dist.is_initialized(), "The default process group is not yet initialized. Please call torch.distributed.init_process_group() first."
            pg = dist.GroupMember.WORLD
        if sharders is None:
            sharders = get_default_sharders()
        return invoke_on_rank_and_broadcast_result(pg, 0, self.plan, module, sharders)


class EmbeddingShardingPlanner(EmbeddingPlannerBase):
    """Finds the sharding plan with the best rated (lowest) step time that fits in memory.

    proposers default to [Greedy, GridSearch(small spaces), Uniform]; partitioner defaults to
    GreedyPerfPartitioner; perf model to the max-over-devices model."""

    def __init__(self, topology: Optional[Topology] = None, batch_size: Optional[int] = None, enumerator: Optional[Enumerator] = None,
                 storage_reservation: Optional[StorageReservation] = None, proposer: Optional[Union[Proposer, List[Proposer]]] = None,
                 partitioner: Optional[Partitioner] = None, performance_model: Optional[PerfModel] = None,
                 stats: Optional[Union[Stats, List[Stats]]] = None, constraints: Optional[Dict[str, ParameterConstraints]] = None, debug: bool = True,
                 callbacks: Optional[List[Callable[[List[ShardingOption]], List[ShardingOption]]]] = None, timeout_seconds: Optional[int] = None,
                 plan_loader=None) -> None:
        super().__init__(topology=topology, batch_size=batch_size, enumerator=enumerator, storage_reservation=storage_reservation, stats=stats,
                         constraints=constraints, debug=debug, callbacks=callbacks, timeout_seconds=timeout_seconds)
        self._partitioner: Partitioner = partitioner if partitioner else GreedyPerfPartitioner()
        if proposer:
            self._proposers: List[Proposer] = [proposer] if not isinstance(proposer, list) else proposer
        else:
            self._proposers = [GridSearchProposer(), GreedyProposer(), GreedyProposer(use_depth=False), UniformProposer()]
        self._perf_model: PerfModel = performance_model if performance_model else NoopPerfModel(topology=self._topology)
        self._num_proposals: int = 0
        self._num_plans: int = 0
        self._best_plan: Optional[List[ShardingOption]] = None
        self._plan_loader = plan_loader

    def hash_planner_context_inputs(self) -> int:
        """Hash of the planner inputs (topology, batch size, search space, reservation, constraints); needs ``plan`` (or the enumerator
        and the reservation) to have run - before that: topology, batch size and constraint names only."""
        from .types import hash_planner_context_inputs

        if getattr(self._enumerator, "last_stored_search_space", None) is not None and self._storage_reservation.last_reserved_topology is not None:
            return hash_planner_context_inputs(self._topology, self._batch_size, self._enumerator, self._storage_reservation, self._constraints)
        parts = [self._topology._hash(), self._batch_size, repr(sorted((self._constraints or {}).keys()))]
        return int(hashlib.sha256(repr(parts).encode()).hexdigest()[:15], 16)

    def hash_planner_context_inputs_str(self) -> str:
        from .types import hash_planner_context_inputs_str

        return hash_planner_context_inputs_str(self._topology, self._batch_size, self._enumerator, self._storage_reservation, self._constraints)

    @property
    def plan_loader(self):
        return self._plan_loader

    @property
    def best_plan(self) -> Optional[List[ShardingOption]]:
        """The chosen sharding options of the last ``plan`` call (what a ``PlanLoader`` stores, keyed by ``storage_hash()``)."""
        return self._best_plan

    def plan(self, module: nn.Module, sharders: Optional[List[ModuleSharder[nn.Module]]] = None) -> ShardingPlan:
        if sharders is None:
            sharders = get_default_sharders()
        self._num_proposals = 0
        self._num_plans = 0
        start_time = time.perf_counter()
        best_plan = None
        lowest_storage = Storage(MAX_SIZE, MAX_SIZE)
        last_planner_error: Optional[PlannerError] = None
        last_proposal: List[ShardingOption] = []
        best_perf_rating = MAX_SIZE

        storage_constraint: Topology = self._storage_reservation.reserve(
            topology=self._topology, batch_size=self._batch_size, module=module, sharders=sharders, constraints=self._constraints)
        search_space = self._enumerator.enumerate(module=module, sharders=sharders)
        if not search_space:
            return ShardingPlan({})
        proposal_cache: Dict[Tuple[int, ...], Tuple[bool, Optional[List[ShardingOption]], Optional[float]]] = {}
        proposers = self._proposers
        if self._plan_loader is not None:
            # a stored plan is only valid for the planner inputs it was computed for
            stored_hash = self._plan_loader.plan_context_hash()
            if stored_hash is not None and stored_hash != self.hash_planner_context_inputs_str():
                raise PlannerError(error_type=PlannerErrorType.PLANNER_INPUT_CONTEXT_MISMATCH,
                                   message="Unable to load, because of planner input mismatch - cannot validate this plan is the best plan for current context.. \n"
                                           f"Planner input context mismatch detected for {self._plan_loader.get_plan_id()} and current planner set up:"
                                           f"\nCurrent planner hash: {self.hash_planner_context_inputs_str()}, Loaded plan hash: {stored_hash}")
            loaded = self._plan_loader.load()
            if loaded is not None:
                best_plan = copy.deepcopy(extract_plan(search_space, loaded))
                best_perf_rating = self._perf_model.rate(plan=best_plan)
                logger.info("Loaded sharding options from storage (plan id %s): skipping the search", self._plan_loader.get_plan_id())
                proposers = []
        for proposer in proposers:
            proposer.load(search_space=search_space, enumerator=self._enumerator)
        for proposer in proposers:
            proposal = proposer.propose()
            while proposal:
                end_time = time.perf_counter()
                if self._timeout_seconds and end_time - start_time > self._timeout_seconds:
                    if best_plan is None:
                        raise PlannerError(error_type=PlannerErrorType.OTHER, message=f"Unable to find a plan within {self._timeout_seconds} s")
                    logger.warning("Planner timed out; returning the best plan found so far")
                    proposal = None
                    break
                proposal_key = tuple(sorted(map(hash, proposal)))
                if proposal_key in proposal_cache:
                    partitionable, plan, perf_rating = proposal_cache[proposal_key]
                    proposer.feedback(partitionable=partitionable, plan=plan, perf_rating=perf_rating, storage_constraint=storage_constraint)
                    proposal = proposer.propose()
                    continue
                self._num_proposals += 1
                try:
                    for cb in self._callbacks:
                        proposal = cb(proposal)
                    plan = self._partitioner.partition(proposal=proposal, storage_constraint=storage_constraint)
                    self._num_plans += 1
                    perf_rating = self._perf_model.rate(plan=plan)
                    if perf_rating < best_perf_rating:
                        best_perf_rating = perf_rating
                        best_plan = copy.deepcopy(plan)
                    proposal_cache[proposal_key] = (True, plan, perf_rating)
                    proposer.feedback(partitionable=True, plan=plan, perf_rating=perf_rating, storage_constraint=storage_constraint)
                except PlannerError as planner_error:
                    last_planner_error = planner_error
                    current_storage = cast(Storage, reduce(lambda x, y: x + y, [shard.storage for option in proposal for shard in option.shards]))
                    if current_storage.hbm < lowest_storage.hbm or (current_storage.hbm == lowest_storage.hbm and current_storage.ddr < lowest_storage.ddr):
                        lowest_storage = current_storage
                    proposal_cache[proposal_key] = (False, proposal, None)
                    proposer.feedback(partitionable=False, plan=proposal, storage_constraint=storage_constraint)
                last_proposal = proposal
                reset_shard_rank(proposal)
                proposal = proposer.propose()
        if best_plan:
            self._best_plan = best_plan
            sharding_plan = to_sharding_plan(best_plan, self._topology)
            end_time = time.perf_counter()
            for stats in self._stats:
                stats.log(sharding_plan=sharding_plan, topology=self._topology, batch_size=self._batch_size, storage_reservation=self._storage_reservation,
                          num_proposals=self._num_proposals, num_plans=self._num_plans, run_time=end_time - start_time, best_plan=best_plan,
                          constraints=self._constraints, sharders=sharders, debug=self._debug)
            return sharding_plan
        global_storage_capacity = reduce(lambda x, y: x + y, [device.storage for device in self._topology.devices])
        global_storage_constraints = reduce(lambda x, y: x + y, [device.storage for device in storage_constraint.devices])
        storage_reservation_solution = (
            f"\\n\\t  Storage reservation is too high? total capacity {storage_repr_in_gb(global_storage_capacity)}, "
            f"after reservations {storage_repr_in_gb(global_storage_constraints)}")
        no_plan_solution = (
            f"Planner evaluated {self._num_proposals} proposals for device {self._topology.compute_device}.\\nPossible solutions:"
            f"\\n  1) Increase the number of devices ({self._topology.world_size})"
            f"\\n  2) Reduce the model size.\\n\\t  Global storage: {round(bytes_to_gb(global_storage_capacity.hbm), 3)} GB HBM & "
            f"{round(bytes_to_gb(global_storage_capacity.ddr), 3)} GB DDR\\n\\t  Available for model parallel: {storage_repr_in_gb(global_storage_constraints)}"
            f"\\n\\t  Requirement for model parallel (smallest proposal): {storage_repr_in_gb(lowest_storage)}"
            f"\\n  3) Reduce local batch size ({self._batch_size})\\n  4) Remove planner constraints that might be reducing the search space or add fused_uvm / row-wise options."
            + storage_reservation_solution)
        if not lowest_storage.fits_in(global_storage_constraints):
            raise PlannerError(error_type=PlannerErrorType.INSUFFICIENT_STORAGE, message="Unable to find a plan for this model because of insufficient storage. \\n" + no_plan_solution)
        raise PlannerError(error_type=PlannerErrorType.STRICT_CONSTRAINTS,
                           message="Unable to find a plan for this model because of a partitioning error: " + str(last_planner_error) + "\\n" + no_plan_solution)

    @property
    def best_plan(self) -> Optional[List[ShardingOption]]:
        return self._best_plan


class HeteroEmbeddingShardingPlanner(EmbeddingShardingPlanner):
    """Planner over heterogeneous device groups (e.g. CPU + GPU inference tiers): one sub-plan per
    ``device_group`` constraint, merged (reference planners.py HeteroEmbeddingShardingPlanner)."""

    def __init__(self, topology_groups: Dict[str, Topology], **kwargs) -> None:
        first = next(iter(topology_groups.values()))
        super().__init__(topology=first, **kwargs)
        self._topology_groups = topology_groups
        self._kwargs = kwargs

    def plan(self, module: nn.Module, sharders: Optional[List[ModuleSharder[nn.Module]]] = None) -> ShardingPlan:
        plans = []
        for group, topo in self._topology_groups.items():
            cons = {k: v for k, v in (self._constraints or {}).items() if v.device_group in (None, group)}
            sub = EmbeddingShardingPlanner(topology=topo, batch_size=self._batch_size, constraints=cons or None, debug=self._debug)
            try:
                plans.append(sub.plan(module, sharders))
            except PlannerError:
                continue
        if not plans:
            raise PlannerError("no device group can host the model")
        return _merge_plans(plans)
