"""2D parallel DMPCollection on 4 CPU ranks (2 sharding groups x 2 replicas), reference test_2d_sharding.py."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _run(ctx, inter_host: bool):
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DMPCollection
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    torch.manual_seed(0)
    W, S, dev = ctx.world_size, 2, ctx.device
    tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=20, feature_names=["f0"]),
              EmbeddingBagConfig(name="t1", embedding_dim=8, num_embeddings=20, feature_names=["f1"])]
    ebc = EmbeddingBagCollection(tables, device=dev)
    apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": 0.5})

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = ebc
            self.lin = torch.nn.Linear(16, 1)

        def forward(self, kjt):
            return self.lin(self.ebc(kjt).values()).sum()

    plan = sp.construct_module_sharding_plan(ebc, {"t0": sp.table_wise(rank=0), "t1": sp.row_wise()}, sharder=EmbeddingBagCollectionSharder(), world_size=S,
                                             local_size=S, device_type=dev.type)
    model = DMPCollection(M(), device=dev, plan=ShardingPlan({"ebc": plan}), world_size=W, sharding_group_size=S, global_pg=dist.group.WORLD,
                          sharders=[EmbeddingBagCollectionSharder()], use_inter_host_allreduce=inter_host)
    assert dist.get_world_size(model.sharding_pg) == S and dist.get_world_size(model.replica_pg) == W // S
    g = torch.Generator().manual_seed(ctx.rank)
    lengths = torch.randint(1, 3, (8,), generator=g)
    kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.randint(0, 20, (int(lengths.sum()),), generator=g), lengths=lengths)
    model(kjt).backward()

    def shard_bytes():
        return torch.cat([t.weights.detach().flatten() for t in model.module.ebc.engine._tbes])

    mine = shard_bytes()
    peers = [torch.empty_like(mine) for _ in range(W // S)]
    dist.all_gather(peers, mine, group=model.replica_pg)
    assert not torch.allclose(peers[0], peers[1]), "replicas saw different batches: shards must have diverged"
    model.sync()
    mine = shard_bytes()
    dist.all_gather(peers, mine, group=model.replica_pg)
    torch.testing.assert_close(peers[0], peers[1])
    # dense weights are DDP-synchronised over the global group
    lw = model.module.lin.weight.grad.clone()
    allw = [torch.empty_like(lw) for _ in range(W)]
    dist.all_gather(allw, lw)
    for a in allw[1:]:
        torch.testing.assert_close(a, allw[0])


@pytest.mark.parametrize("inter_host", [False, True])
def test_dmp_collection_sync(inter_host):
    run_multi_process(_run, world_size=4, backend="gloo", inter_host=inter_host)


def _run_fully_sharded(ctx):
    """FULLY_SHARDED == DEFAULT + a weight sync between every forward and its backward; storage is released in between."""
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DMPCollection
    from torchrec_b200.parallel.types import ShardingPlan, ShardingStrategy
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, S, dev = ctx.world_size, 2, ctx.device

    def build(strategy):
        torch.manual_seed(0)
        tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=21, feature_names=["f0"]),
                  EmbeddingBagConfig(name="t1", embedding_dim=8, num_embeddings=20, feature_names=["f1"])]
        ebc = EmbeddingBagCollection(tables, device=dev)
        apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": 0.5})

        class M(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.ebc = ebc
                self.lin = torch.nn.Linear(16, 1)

            def forward(self, kjt):
                return self.lin(self.ebc(kjt).values()).sum()

        plan = sp.construct_module_sharding_plan(ebc, {"t0": sp.table_wise(rank=0), "t1": sp.row_wise()}, sharder=EmbeddingBagCollectionSharder(),
                                                 world_size=S, local_size=S, device_type=dev.type)
        return DMPCollection(M(), device=dev, plan=ShardingPlan({"ebc": plan}), world_size=W, sharding_group_size=S, global_pg=dist.group.WORLD,
                             sharders=[EmbeddingBagCollectionSharder()], sharding_strategy=strategy)

    fs, ref = build(ShardingStrategy.FULLY_SHARDED), build(ShardingStrategy.DEFAULT)
    tbes = fs.module.ebc.engine._tbes
    assert len(fs._fully_sharded) == len(tbes) > 0 and not ref._fully_sharded
    g = torch.Generator().manual_seed(100 + ctx.rank)
    for step in range(3):
        lengths = torch.randint(1, 3, (8,), generator=g)
        kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.randint(0, 20, (int(lengths.sum()),), generator=g), lengths=lengths)
        loss = fs(kjt)
        for t in tbes:  # between forward and backward only 1/R of the averaged buffer is resident
            assert t._fs.sharded and t.weights.data.untyped_storage().nbytes() == 0 and t._fs.shard.numel() * 2 >= t._fs.numel
        loss.backward()
        for t in tbes:
            assert not t._fs.sharded and t.weights.data.untyped_storage().nbytes() == t._fs.nbytes
        loss_ref = ref(kjt)
        ref.sync(include_optimizer_state=False)
        loss_ref.backward()
        torch.testing.assert_close(loss.detach(), loss_ref.detach())
    for a, b in zip(tbes, ref.module.ebc.engine._tbes):
        torch.testing.assert_close(a.weights.detach(), b.weights.detach())
    # a forward without backward leaves the shards split; state_dict() / await_rs_awaitables() make them whole
    fs(kjt)
    assert all(t._fs.sharded for t in tbes)
    sd = fs.state_dict()
    assert not any(t._fs.sharded for t in tbes) and len(sd) > 0
    fs.eval()
    with torch.no_grad():
        fs(kjt)
    assert not any(t._fs.sharded for t in tbes)


def test_dmp_collection_fully_sharded():
    run_multi_process(_run_fully_sharded, world_size=4, backend="gloo")
