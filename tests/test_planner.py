"""Planner unit tests (no process group needed; mirrors the reference's planner/tests strategy)."""
import pytest
import torch

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
from torchrec_b200.parallel.planner import (
    DynamicProgrammingProposer,
    EmbeddingEnumerator,
    EmbeddingShardingPlanner,
    GreedyPerfPartitioner,
    MemoryBalancedPartitioner,
    ParameterConstraints,
    PlannerError,
    Topology,
)
from torchrec_b200.parallel.types import ShardingType


class Model(torch.nn.Module):
    def __init__(self, tables):
        super().__init__()
        self.sparse = EmbeddingBagCollection(tables, device=torch.device("meta"))


def _tables(n=6, rows=100_000, dim=128):
    return [EmbeddingBagConfig(name=f"t{i}", embedding_dim=dim, num_embeddings=rows * (i + 1), feature_names=[f"f{i}"]) for i in range(n)]


def test_enumerator_lists_all_sharding_types_single_domain():
    topo = Topology(world_size=8, compute_device="cuda")
    opts = EmbeddingEnumerator(topo, batch_size=512).enumerate(Model(_tables(2, dim=256)), [EmbeddingBagCollectionSharder()])
    types = {o.sharding_type for o in opts if o.name == "t0"}
    # single NVLink domain: TWRW / TWCW / GRID collapse onto RW / CW
    assert {"data_parallel", "table_wise", "row_wise", "column_wise"} <= types
    assert "table_row_wise" not in types and "grid_shard" not in types
    for o in opts:
        for s in o.shards:
            assert s.storage is not None and s.perf is not None and s.perf.total >= 0
    rw = [o for o in opts if o.name == "t0" and o.sharding_type == "row_wise"][0]
    assert sum(s.size[0] for s in rw.shards) == 100_000


def test_enumerator_multi_host_adds_hierarchical_types():
    topo = Topology(world_size=16, local_world_size=8, compute_device="cuda")
    opts = EmbeddingEnumerator(topo, batch_size=512).enumerate(Model(_tables(1, dim=512)), [EmbeddingBagCollectionSharder()])
    types = {o.sharding_type for o in opts}
    assert {"table_row_wise", "table_column_wise", "grid_shard"} <= types


def test_planner_balances_and_respects_constraints():
    topo = Topology(world_size=4, compute_device="cuda")
    constraints = {"t0": ParameterConstraints(sharding_types=["row_wise"]), "t1": ParameterConstraints(sharding_types=["column_wise"], min_partition=64)}
    planner = EmbeddingShardingPlanner(topology=topo, batch_size=1024, constraints=constraints)
    plan = planner.plan(Model(_tables()), [EmbeddingBagCollectionSharder()])
    mp = plan.plan["sparse"]
    assert mp["t0"].sharding_type == "row_wise" and len(mp["t0"].ranks) == 4
    assert mp["t1"].sharding_type == "column_wise" and len(mp["t1"].sharding_spec.shards) == 2
    used = set(r for ps in mp.values() for r in ps.ranks)
    assert used == {0, 1, 2, 3}
    assert any("Planner Statistics" in l for l in planner._stats[0]._stats_table)


def test_planner_insufficient_storage_raises():
    topo = Topology(world_size=2, compute_device="cuda", hbm_cap=64 * 1024 * 1024)
    cons = {f"t{i}": ParameterConstraints(compute_kernels=["fused"]) for i in range(4)}  # no host-offloaded (UVM) options
    planner = EmbeddingShardingPlanner(topology=topo, batch_size=512, constraints=cons)
    with pytest.raises(PlannerError):
        planner.plan(Model(_tables(4, rows=10_000_000)), [EmbeddingBagCollectionSharder()])


def test_big_table_forces_row_wise_when_it_does_not_fit_one_gpu():
    topo = Topology(world_size=8, compute_device="cuda", hbm_cap=8 * 1024**3)
    tables = [EmbeddingBagConfig(name="big", embedding_dim=128, num_embeddings=40_000_000, feature_names=["f"])]  # 20 GB fp32
    cons = {"big": ParameterConstraints(compute_kernels=["fused"])}
    plan = EmbeddingShardingPlanner(topology=topo, batch_size=512, constraints=cons).plan(Model(tables), [EmbeddingBagCollectionSharder()])
    assert plan.plan["sparse"]["big"].sharding_type in ("row_wise", "column_wise")


def test_partitioners_and_dp_proposer():
    topo = Topology(world_size=4, compute_device="cuda")
    m = Model(_tables())
    for part in (GreedyPerfPartitioner(), MemoryBalancedPartitioner()):
        plan = EmbeddingShardingPlanner(topology=topo, batch_size=512, partitioner=part, proposer=DynamicProgrammingProposer()).plan(m, [EmbeddingBagCollectionSharder()])
        assert set(plan.plan["sparse"].keys()) == {f"t{i}" for i in range(6)}


def test_sharding_route_facade():
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.sharding import all_routes, describe, units_for

    assert {r.sharding_type for r in all_routes()} == {"data_parallel", "table_wise", "column_wise", "table_column_wise", "row_wise", "table_row_wise", "grid_shard"}
    assert "sum" in describe("row_wise").combine and describe("column_wise").reference_module.endswith("cw_sharding.py")
    cfg = EmbeddingBagConfig(name="t", embedding_dim=16, num_embeddings=100, feature_names=["f"])
    ebc = EmbeddingBagCollection([cfg], device="meta")
    plan = sp.construct_module_sharding_plan(ebc, {"t": sp.row_wise()}, sharder=EmbeddingBagCollectionSharder(), world_size=4, local_size=4, device_type="cpu")
    shards = units_for(0, cfg, plan["t"])
    assert [s.rank for s in shards] == [0, 1, 2, 3] and sum(s.rows for s in shards) == 100 and all(s.cols == 16 for s in shards)
