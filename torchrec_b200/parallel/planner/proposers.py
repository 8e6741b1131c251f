"""Plan proposers (reference planner/proposers.py:34-983): Greedy, Uniform, GridSearch, DynamicProgramming,
EmbeddingOffloadScaleup."""
from __future__ import annotations

import copy
import itertools
import logging
from decimal import Decimal
from typing import Callable, Dict, List, Optional, Set, Tuple, Union, cast

import torch

from ..embedding_types import EmbeddingComputeKernel
from ..types import CacheAlgorithm, CacheParams
from .types import Enumerator, Perf, Proposer, ShardingOption, Storage, Topology
from .utils import BinarySearchPredicate, bytes_to_gb, prod

logger = logging.getLogger(__name__)

MAX_PROPOSALS: int = int(1e4)


class GreedyProposer(Proposer):
    """Per table, options sorted by perf; proposes the best of each, then walks down the option list of the
    table that is blamed for the failure (largest HBM user) or the largest perf contributor."""

    def __init__(self, use_depth: bool = True, threshold: Optional[int] = None) -> None:
        self._use_depth = use_depth
        self._threshold = threshold if threshold else 0
        self._sharding_options_by_fqn: Dict[str, List[ShardingOption]] = {}
        self._current_proposal: Dict[str, int] = {}
        self._best_perf_rating: float = float("inf")
        self._num_inferior_perf: int = 0

    def load(self, search_space: List[ShardingOption], enumerator: Optional[Enumerator] = None) -> None:
        self._reset()
        for so in search_space:
            self._sharding_options_by_fqn.setdefault(so.fqn, []).append(so)
        for sharding_options in self._sharding_options_by_fqn.values():
            sharding_options.sort(key=lambda x: _sharding_option_score(x, self._use_depth))
        self._current_proposal = {fqn: 0 for fqn in self._sharding_options_by_fqn.keys()}

    def _reset(self) -> None:
        self._sharding_options_by_fqn = {}
        self._current_proposal = {}

    def propose(self) -> Optional[List[ShardingOption]]:
        if self._current_proposal:
            return [self._sharding_options_by_fqn[fqn][index] for fqn, index in self._current_proposal.items()]
        return None

    def feedback(self, partitionable: bool, plan: Optional[List[ShardingOption]] = None, perf_rating: Optional[float] = None,
                 storage_constraint: Optional[Topology] = None) -> None:
        if self._threshold and perf_rating is not None:
            self._num_inferior_perf += 1
            if perf_rating < self._best_perf_rating:
                self._best_perf_rating = perf_rating
                self._num_inferior_perf = 0
            if self._num_inferior_perf > self._threshold:
                self._current_proposal = {}
                return
        # move the table with the largest storage footprint to its next option
        largest_fqn = None
        largest_storage: Tuple[float, float, float, float] = (0, 0, 0, 0)
        for fqn, options in self._sharding_options_by_fqn.items():
            index = self._current_proposal[fqn]
            if index + 1 < len(options):
                so = options[index]
                current_storage = (
                    max(shard.storage.hbm for shard in so.shards),  # type: ignore[union-attr]
                    sum(shard.storage.hbm for shard in so.shards),  # type: ignore[union-attr]
                    max(shard.storage.ddr for shard in so.shards),  # type: ignore[union-attr]
                    sum(shard.storage.ddr for shard in so.shards),  # type: ignore[union-attr]
                )
                if current_storage > largest_storage:
                    largest_fqn = fqn
                    largest_storage = current_storage
        if largest_fqn is not None:
            self._current_proposal[largest_fqn] += 1
        else:
            self._current_proposal = {}


class UniformProposer(Proposer):
    """Proposes plans where every table uses the same sharding type."""

    def __init__(self, use_depth: bool = True) -> None:
        self._use_depth = use_depth
        self._grouped_sharding_options: List[List[ShardingOption]] = []
        self._proposal_index: int = 0

    def load(self, search_space: List[ShardingOption], enumerator: Optional[Enumerator] = None) -> None:
        self._reset()
        all_fqns = set()
        by_type_and_fqn: Dict[str, Dict[str, List[ShardingOption]]] = {}
        for so in search_space:
            by_type_and_fqn.setdefault(so.sharding_type, {}).setdefault(so.fqn, []).append(so)
            all_fqns.add(so.fqn)
        for by_fqn in by_type_and_fqn.values():
            if by_fqn.keys() == all_fqns:
                self._grouped_sharding_options.append([
                    sorted(by_fqn[fqn], key=lambda x: _sharding_option_score(x, self._use_depth))[0] for fqn in all_fqns])

    def _reset(self) -> None:
        self._grouped_sharding_options = []
        self._proposal_index = 0

    def propose(self) -> Optional[List[ShardingOption]]:
        if self._proposal_index < len(self._grouped_sharding_options):
            return self._grouped_sharding_options[self._proposal_index]
        return None

    def feedback(self, partitionable: bool, plan=None, perf_rating=None, storage_constraint=None) -> None:
        self._proposal_index += 1


class GridSearchProposer(Proposer):
    """Exhaustive search over all combinations (capped at ``max_proposals``)."""

    def __init__(self, max_proposals: int = MAX_PROPOSALS) -> None:
        self._max_proposals = max_proposals
        self._sharding_options_by_fqn: Dict[str, List[ShardingOption]] = {}
        self._proposal_index: int = 0
        self._proposals: List[List[int]] = []

    def load(self, search_space: List[ShardingOption], enumerator: Optional[Enumerator] = None) -> None:
        self._reset()
        for so in search_space:
            self._sharding_options_by_fqn.setdefault(so.fqn, []).append(so)
        total_proposals = prod([len(x) for x in self._sharding_options_by_fqn.values()])
        if total_proposals > self._max_proposals:
            total_proposals = "{:.2e}".format(Decimal(total_proposals))
            logger.info(f"Skipping grid search proposer as there are too many proposals.\nTotal proposals to search: {total_proposals}\n"
                        f"Max proposals allowed: {self._max_proposals}\n")
            return
        sharding_options_by_fqn_indices = [range(len(sos)) for sos in self._sharding_options_by_fqn.values()]
        self._proposals = list(itertools.product(*sharding_options_by_fqn_indices))  # type: ignore[arg-type]

    def _reset(self) -> None:
        self._sharding_options_by_fqn = {}
        self._proposal_index = 0
        self._proposals = []

    def propose(self) -> Optional[List[ShardingOption]]:
        if self._proposals and self._proposal_index < len(self._proposals):
            proposal_indices = self._proposals[self._proposal_index]
            return [sos[index] for index, sos in zip(proposal_indices, self._sharding_options_by_fqn.values())]
        return None

    def feedback(self, partitionable: bool, plan=None, perf_rating=None, storage_constraint=None) -> None:
        self._proposal_index += 1


class DynamicProgrammingProposer(Proposer):
    """Knapsack-style search: minimise total perf subject to a total-HBM budget, HBM discretised into bins.
    Proposes plans for decreasing HBM budgets (reference proposers.py:287-470)."""

    def __init__(self, hbm_bins_per_device: int = 100) -> None:
        self._inited = False
        self._hbm_bins_per_device = max(hbm_bins_per_device, 1)
        self._sharding_options: List[List[ShardingOption]] = []
        self._proposal_list: List[List[int]] = []
        self._current_proposal: int = -1

    def load(self, search_space: List[ShardingOption], enumerator: Optional[Enumerator] = None) -> None:
        self._reset()
        by_fqn: Dict[str, List[ShardingOption]] = {}
        for so in search_space:
            by_fqn.setdefault(so.fqn, []).append(so)
        self._sharding_options = list(by_fqn.values())

    def _reset(self) -> None:
        self._sharding_options = []
        self._proposal_list = []
        self._current_proposal = -1
        self._inited = False

    def propose(self) -> Optional[List[ShardingOption]]:
        if not self._inited:
            return [sorted(p, key=lambda x: _sharding_option_score(x))[0] for p in self._sharding_options]
        if 0 <= self._current_proposal < len(self._proposal_list):
            idx = self._proposal_list[self._current_proposal]
            return [self._sharding_options[i][j] for i, j in enumerate(idx)]
        return None

    def feedback(self, partitionable: bool, plan=None, perf_rating=None, storage_constraint: Optional[Topology] = None) -> None:
        if not self._inited:
            self._inited = True
            table_count = len(self._sharding_options)
            self._proposal_list = []
            if storage_constraint is None or table_count == 0:
                self._current_proposal = 0
                return
            hbm_total = sum(x.storage.hbm for x in storage_constraint.devices)
            bin_count = self._hbm_bins_per_device * len(storage_constraint.devices)
            bin_size = float(hbm_total) / bin_count if hbm_total > 0 else 1.0
            INF = float("inf")
            dp = [[(INF, INF)] * bin_count for _ in range(table_count)]  # (perf, hbm)
            backtrack = [[(-1, -1)] * bin_count for _ in range(table_count)]
            hbm_by_fqn = [[INF] * len(opts) for opts in self._sharding_options]
            perf_by_fqn = [[INF] * len(opts) for opts in self._sharding_options]
            for ti, opts in enumerate(self._sharding_options):
                for oi, so in enumerate(opts):
                    hbm_by_fqn[ti][oi] = _bytes_to_float_bin(so.total_storage.hbm, bin_size)
                    perf_by_fqn[ti][oi] = so.total_perf
            for oi in range(len(self._sharding_options[0])):
                hbm, perf = hbm_by_fqn[0][oi], perf_by_fqn[0][oi]
                b = int(hbm)
                if b < bin_count and dp[0][b][0] > perf:
                    dp[0][b] = (perf, hbm)
                    backtrack[0][b] = (oi, -1)
            for ti in range(1, table_count):
                for oi in range(len(self._sharding_options[ti])):
                    hbm, perf = hbm_by_fqn[ti][oi], perf_by_fqn[ti][oi]
                    for pb in range(bin_count):
                        prev_perf, prev_hbm = dp[ti - 1][pb]
                        if prev_perf == INF:
                            continue
                        nb = int(prev_hbm + hbm)
                        if nb < bin_count and dp[ti][nb][0] > prev_perf + perf:
                            dp[ti][nb] = (prev_perf + perf, prev_hbm + hbm)
                            backtrack[ti][nb] = (oi, pb)
            min_perf = INF
            for b in range(bin_count - 1, -1, -1):
                cur_perf = dp[table_count - 1][b][0]
                if cur_perf < min_perf:
                    min_perf = cur_perf
                    idxs = [-1] * table_count
                    cb = b
                    for ti in range(table_count - 1, -1, -1):
                        idxs[ti], cb = backtrack[ti][cb]
                    self._proposal_list.append(idxs)
            self._proposal_list.reverse()
            self._current_proposal = 0
        else:
            self._current_proposal += 1


def _bytes_to_float_bin(num_bytes: Union[float, int], bin_size: float) -> float:
    return float(num_bytes) / bin_size


class EmbeddingOffloadScaleupProposer(Proposer):
    """For host-offloaded (fused_uvm_caching) tables: once a feasible plan is found, spend the left-over HBM on
    larger cache load factors (binary search over the budget). Reference proposers.py:471-983."""

    def __init__(self, use_depth: bool = True) -> None:
        self.use_depth = use_depth
        self.enumerator: Optional[Enumerator] = None
        self.starting_proposal: List[ShardingOption] = []
        self.proposal: Optional[List[ShardingOption]] = None
        self.search: Optional[BinarySearchPredicate] = None
        self.best_perf_rating = float("inf")

    def load(self, search_space: List[ShardingOption], enumerator: Optional[Enumerator] = None) -> None:
        self.enumerator = enumerator
        by_fqn: Dict[str, List[ShardingOption]] = {}
        for so in search_space:
            by_fqn.setdefault(so.fqn, []).append(so)
        for sos in by_fqn.values():
            sos.sort(key=lambda x: _sharding_option_score(x, self.use_depth))
        proposal = [sos[0] for sos in by_fqn.values()]
        self.starting_proposal = copy.deepcopy(proposal)
        self.proposal = copy.deepcopy(self.starting_proposal)

    def propose(self) -> Optional[List[ShardingOption]]:
        return self.proposal

    def feedback(self, partitionable: bool, plan: Optional[List[ShardingOption]] = None, perf_rating: Optional[float] = None,
                 storage_constraint: Optional[Topology] = None) -> None:
        if not self.enumerator or plan is None or storage_constraint is None:
            self.proposal = None
            return
        cacheable = [so for so in plan if so.compute_kernel == EmbeddingComputeKernel.FUSED_UVM_CACHING.value]
        if not cacheable:
            self.proposal = None
            return
        hbm_available = sum(d.storage.hbm for d in storage_constraint.devices)
        hbm_used = sum(so.total_storage.hbm for so in plan)
        if self.search is None:
            if not partitionable or hbm_used >= hbm_available:
                self.proposal = None
                return
            self.search = BinarySearchPredicate(0, hbm_available - hbm_used, max(1, (hbm_available - hbm_used) // 32))
        budget = self.search.next(partitionable)
        if budget is None:
            self.proposal = None
            return
        self.proposal = copy.deepcopy(self.starting_proposal)
        scaled = [so for so in self.proposal if so.compute_kernel == EmbeddingComputeKernel.FUSED_UVM_CACHING.value]
        total_ddr = sum(so.total_storage.ddr for so in scaled) or 1
        for so in scaled:
            share = budget * so.total_storage.ddr / total_ddr
            clf = min(1.0, (so.cache_load_factor or 0.2) + share / max(so.total_storage.ddr, 1))
            so.cache_params = CacheParams(algorithm=so.cache_params.algorithm if so.cache_params else None, load_factor=clf,
                                          reserved_memory=so.cache_params.reserved_memory if so.cache_params else None,
                                          precision=so.cache_params.precision if so.cache_params else None,
                                          prefetch_pipeline=so.cache_params.prefetch_pipeline if so.cache_params else None,
                                          stats=so.cache_params.stats if so.cache_params else None)
            if clf >= 1.0:
                so.compute_kernel = EmbeddingComputeKernel.FUSED.value
        self.enumerator.populate_estimates(scaled)


def _sharding_option_score(sharding_option: ShardingOption, use_depth: bool = True) -> float:
    return max(cast(Perf, shard.perf).total for shard in sharding_option.shards) if use_depth else sum(
        cast(Perf, shard.perf).total for shard in sharding_option.shards)


def proposers_to_proposals_list(proposers_list: List[Proposer], search_space: List[ShardingOption]) -> List[List[ShardingOption]]:
    """Run proposers (without feedback-driven search) and dedupe their proposals."""
    proposals_list: List[List[ShardingOption]] = []
    proposal_cache: Set[Tuple[int, ...]] = set()
    for proposer in proposers_list:
        proposer.load(search_space=search_space)
        proposal = proposer.propose()
        while proposal:
            proposal_key = tuple(sorted(map(hash, proposal)))
            proposer.feedback(partitionable=True)
            if proposal_key in proposal_cache:
                proposal = proposer.propose()
                continue
            proposals_list.append(proposal)
            proposal_cache.add(proposal_key)
            proposal = proposer.propose()
    return proposals_list
