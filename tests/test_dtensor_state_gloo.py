"""``ShardingEnv(output_dtensor=True)``: state dicts carry DTensors over LocalShardsWrapper (several local shards for column-wise tables)
with the reference's placements, optimizer state included; they load back in place and into a differently initialised model."""
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _run(ctx):
    import torch.distributed as dist
    from torch.distributed.tensor import DTensor, Replicate, Shard

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.shards_wrapper import LocalShardsWrapper
    from torchrec_b200.parallel.types import ShardingEnv, ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    W = ctx.world_size
    tables = [EmbeddingBagConfig(name="tw", embedding_dim=8, num_embeddings=20, feature_names=["a"]),
              EmbeddingBagConfig(name="rw", embedding_dim=8, num_embeddings=30, feature_names=["b"]),
              EmbeddingBagConfig(name="cw", embedding_dim=16, num_embeddings=10, feature_names=["c"])]
    gens = {"tw": sp.table_wise(rank=1), "rw": sp.row_wise(), "cw": sp.column_wise(ranks=[0, 1, 0, 1])}

    def build(seed):
        torch.manual_seed(seed)
        ebc = EmbeddingBagCollection(tables)
        apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.1})
        plan = sp.construct_module_sharding_plan(ebc, gens, sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type="cpu")
        env = ShardingEnv.from_process_group(dist.group.WORLD, output_dtensor=True)
        return DistributedModelParallel(ebc, env=env, device=torch.device("cpu"), plan=ShardingPlan({"": plan}), sharders=[EmbeddingBagCollectionSharder()])

    a = build(0)
    g = torch.Generator().manual_seed(3 + ctx.rank)
    kjt = KeyedJaggedTensor(keys=["a", "b", "c"], values=torch.cat([torch.randint(0, 20, (4,), generator=g), torch.randint(0, 30, (4,), generator=g),
                                                                     torch.randint(0, 10, (4,), generator=g)]), lengths=torch.ones(12, dtype=torch.int64))
    a(kjt).values().sum().backward()
    sd = a.state_dict()
    want = {"embedding_bags.tw.weight": ((Replicate(),), (20, 8)), "embedding_bags.rw.weight": ((Shard(0),), (30, 8)), "embedding_bags.cw.weight": ((Shard(1),), (10, 16))}
    for k, (placements, shape) in want.items():
        t = sd[k]
        assert isinstance(t, DTensor), (k, type(t))
        assert tuple(t.placements) == placements and tuple(t.shape) == shape
        loc = t.to_local()
        assert isinstance(loc, LocalShardsWrapper)
    cw = sd["embedding_bags.cw.weight"].to_local()
    assert len(cw.local_shards()) == 2 and [o[1] for o in cw.local_offsets()] == ([0, 8] if ctx.rank == 0 else [4, 12])
    assert len(sd["embedding_bags.tw.weight"].to_local().local_shards()) == (1 if ctx.rank == 1 else 0)
    osd = a.fused_optimizer.state_dict()["state"]
    m1 = osd["embedding_bags.rw.weight"]["rw.momentum1"]
    assert isinstance(m1, DTensor) and tuple(m1.shape) == (30,) and tuple(m1.placements) == (Shard(0),)
    # round trip into a differently initialised model (same plan): weights + optimizer state
    b = build(99)
    b(kjt).values().sum().backward()  # materialise / perturb b's state
    b.load_state_dict(sd)
    b.fused_optimizer.load_state_dict(a.fused_optimizer.state_dict())
    torch.testing.assert_close(b(kjt).values(), a(kjt).values())
    a(kjt).values().sum().backward()
    b(kjt).values().sum().backward()
    torch.testing.assert_close(b(kjt).values(), a(kjt).values(), rtol=1e-5, atol=1e-6)


def test_dtensor_state_dict_layout_and_round_trip():
    run_multi_process(_run, world_size=2, backend="gloo")
