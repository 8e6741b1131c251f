"""Golden-model test of ShardedEmbeddingBagCollection on 2 CPU ranks (gloo): a sharded model and
an unsharded model built from the same tables take one train step; outputs and updated weights
must match (methodology of the reference's sharding_single_rank_test, test_sharding.py:775-1065)."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _tables(weighted=False):
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, PoolingType

    return [
        EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=40, feature_names=["f0"]),
        EmbeddingBagConfig(name="t1", embedding_dim=16, num_embeddings=30, feature_names=["f1", "f2"]),
        EmbeddingBagConfig(name="t2", embedding_dim=8, num_embeddings=20, feature_names=["f3"], pooling=PoolingType.SUM if weighted else PoolingType.MEAN),
        EmbeddingBagConfig(name="t3", embedding_dim=8, num_embeddings=25, feature_names=["f4"]),
    ]


def _make_batch(rank, B, weighted):
    from torchrec_b200.sparse import KeyedJaggedTensor

    g = torch.Generator().manual_seed(100 + rank)
    keys = ["f0", "f1", "f2", "f3", "f4"]
    hashes = [40, 30, 30, 20, 25]
    lengths = torch.randint(0, 4, (len(keys) * B,), generator=g)
    vals = []
    for i, h in enumerate(hashes):
        n = int(lengths[i * B : (i + 1) * B].sum())
        vals.append(torch.randint(0, h, (n,), generator=g))
    values = torch.cat(vals)
    w = torch.rand(values.numel(), generator=g) if weighted else None
    return KeyedJaggedTensor(keys=keys, values=values, lengths=lengths, weights=w)


def _run(ctx, sharding: str, weighted: bool):
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.parallel.comm_ops import set_gradient_division

    set_gradient_division(False)
    torch.manual_seed(0)
    W, B = ctx.world_size, 6
    dev = ctx.device
    gold = EmbeddingBagCollection(_tables(weighted), is_weighted=weighted, device=dev)
    local = EmbeddingBagCollection(_tables(weighted), is_weighted=weighted, device=dev)
    local.load_state_dict(gold.state_dict())
    # row-wise adagrad over column shards normalises per shard (as in the reference), so the
    # column-sharded configurations are checked with SGD
    use_sgd = sharding in ("cw", "mixed", "twrw")
    if use_sgd:
        apply_optimizer_in_backward(torch.optim.SGD, local.parameters(), {"lr": 0.1})
    else:
        apply_optimizer_in_backward(RowWiseAdagrad, local.parameters(), {"lr": 0.1, "eps": 1e-8})
    gens = {
        "tw": {"t0": sp.table_wise(rank=0), "t1": sp.table_wise(rank=1), "t2": sp.table_wise(rank=0), "t3": sp.table_wise(rank=1)},
        "rw": {n: sp.row_wise() for n in ["t0", "t1", "t2", "t3"]},
        "cw": {"t0": sp.column_wise(ranks=[0, 1]), "t1": sp.column_wise(ranks=[1, 0]), "t2": sp.table_wise(rank=1), "t3": sp.column_wise(ranks=[0, 1])},
        "mixed": {"t0": sp.row_wise(), "t1": sp.column_wise(ranks=[0, 1]), "t2": sp.row_wise(), "t3": sp.data_parallel()},
        "twrw": {"t0": sp.table_row_wise(host_index=0), "t1": sp.table_row_wise(host_index=0), "t2": sp.table_wise(rank=1), "t3": sp.grid_shard(host_indexes=[0])},
    }[sharding]
    plan = sp.construct_module_sharding_plan(local, gens, sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type=dev.type)

    class Wrap(torch.nn.Module):
        def __init__(self, ebc):
            super().__init__()
            self.ebc = ebc

        def forward(self, kjt):
            return self.ebc(kjt).values()

    model = DistributedModelParallel(Wrap(local), device=dev, plan=ShardingPlan({"ebc": plan}),
                                     sharders=[EmbeddingBagCollectionSharder()])
    dense_params = [p for n, p in model.named_parameters() if p.requires_grad]
    dense_opt = torch.optim.SGD(dense_params, lr=0.1) if dense_params else None
    gold_opt = torch.optim.SGD(gold.parameters(), lr=0.1) if use_sgd else RowWiseAdagrad(gold.parameters(), lr=0.1, eps=1e-8)
    for step in range(2):
        batches = [_make_batch(r + 10 * step, B, weighted).to(dev) for r in range(W)]
        out = model(batches[ctx.rank])
        # golden on the global batch
        gouts = [gold(b).values() for b in batches]
        torch.testing.assert_close(out.float(), gouts[ctx.rank], rtol=1e-5, atol=1e-5)
        proj = torch.linspace(0.5, 1.5, out.shape[1], device=dev)
        (out * proj).sum().backward()
        if dense_opt is not None:
            # DDP averages dense grads over ranks; the golden sums the per-rank losses -> scale
            for p in dense_params:
                p.grad.mul_(W)
            dense_opt.step()
            dense_opt.zero_grad()
        gold_opt.zero_grad()
        sum((o * proj).sum() for o in gouts).backward()
        if "t3" in [n for n in ["t3"] if sharding == "mixed"]:
            # data-parallel table trained with plain SGD in the sharded model
            with torch.no_grad():
                w3 = gold.embedding_bags["t3"].weight
                g3 = w3.grad.clone()
                w3.grad = None
                gold_opt.step()
                w3.sub_(0.1 * g3)
        else:
            gold_opt.step()
    # compare weights: gather full tables from the sharded state dict
    sd = model.state_dict()
    for name in ["t0", "t1", "t2", "t3"]:
        st = sd[f"ebc.embedding_bags.{name}.weight"]
        ref = gold.embedding_bags[name].weight.detach()
        if hasattr(st, "local_shards"):
            for sh in st.local_shards():
                o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
                torch.testing.assert_close(sh.tensor, ref[o[0] : o[0] + s[0], o[1] : o[1] + s[1]], rtol=1e-4, atol=1e-5)
        else:
            torch.testing.assert_close(st, ref, rtol=1e-4, atol=1e-5)
    # fused optimizer state is keyed by table
    fo = model.fused_optimizer.state_dict()["state"]
    assert all(k.startswith("ebc.embedding_bags.") for k in fo.keys()), list(fo.keys())
    if sharding == "rw":
        assert "ebc.embedding_bags.t0.weight" in fo and "t0.momentum1" in fo["ebc.embedding_bags.t0.weight"]


@pytest.mark.parametrize("sharding", ["tw", "rw", "cw", "mixed", "twrw"])
@pytest.mark.parametrize("weighted", [False, True])
def test_sharded_ebc_matches_unsharded(sharding, weighted):
    run_multi_process(_run, world_size=2, backend="gloo", sharding=sharding, weighted=weighted)


def _rank_without_shards(ctx):
    """A module whose only table lives on ONE rank: the other rank holds no shard but must still join the backward collectives."""
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embedding import EmbeddingCollectionSharder
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingEnv, ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, dev = ctx.world_size, ctx.device
    torch.manual_seed(0)
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="t", embedding_dim=8, num_embeddings=30, feature_names=["f"])])
    ec = EmbeddingCollection([EmbeddingConfig(name="s", embedding_dim=8, num_embeddings=30, feature_names=["g"])])
    for m in (ebc, ec):
        apply_optimizer_in_backward(torch.optim.SGD, m.parameters(), {"lr": 0.1})

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc, self.ec = ebc, ec
            self.lin = torch.nn.Linear(8, 1)

        def forward(self, kjt):
            pooled = self.ebc(kjt).values()
            seq = self.ec(kjt)["g"].values()
            return self.lin(pooled).sum() + seq.sum()

    plan = ShardingPlan({
        "ebc": sp.construct_module_sharding_plan(ebc, {"t": sp.table_wise(rank=W - 1)}, sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type=dev.type),
        "ec": sp.construct_module_sharding_plan(ec, {"s": sp.table_wise(rank=0)}, sharder=EmbeddingCollectionSharder(), world_size=W, local_size=W, device_type=dev.type),
    })
    dmp = DistributedModelParallel(M(), env=ShardingEnv.from_process_group(dist.group.WORLD), device=dev, plan=plan,
                                   sharders=[EmbeddingBagCollectionSharder(), EmbeddingCollectionSharder()])
    g = torch.Generator().manual_seed(ctx.rank)
    for _ in range(2):
        lengths = torch.randint(1, 3, (2 * 4,), generator=g)
        kjt = KeyedJaggedTensor(keys=["f", "g"], values=torch.randint(0, 30, (int(lengths.sum()),), generator=g), lengths=lengths)
        dmp(kjt).backward()      # hung before: the shard-less rank skipped the output dist's backward all-to-all


def test_rank_without_shards_joins_backward_collectives():
    run_multi_process(_rank_without_shards, world_size=2, backend="gloo")


def _run_embedding_module_interface(ctx):
    """ShardedEmbeddingModule surface of the sharded bags: sharding scheme by type, post-lookup tracker callback with the ids this rank
    looks up, output-dist tracker registration."""
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embedding_types import ShardedEmbeddingModule
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=20, feature_names=["f0", "f0b"]), EmbeddingBagConfig(name="t1", embedding_dim=8, num_embeddings=30, feature_names=["f1"])]
    ebc = EmbeddingBagCollection(tables)
    sharder = EmbeddingBagCollectionSharder()
    plan = sp.construct_module_sharding_plan(ebc, {"t0": sp.table_wise(rank=1), "t1": sp.row_wise()}, sharder=sharder, world_size=2, local_size=2, device_type="cpu")

    class W(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = ebc

        def forward(self, k):
            return self.ebc(k).values()

    dmp = DistributedModelParallel(W(), device=torch.device("cpu"), plan=ShardingPlan({"ebc": plan}), sharders=[sharder])
    sebc = dmp.module.ebc
    assert isinstance(sebc, ShardedEmbeddingModule) and sebc.unsharded_module_type is EmbeddingBagCollection
    sh = sebc.shardings
    assert set(sh) == {"table_wise", "row_wise"}
    assert sh["table_wise"].feature_names() == ["f0", "f0b"] and sh["table_wise"].feature_names_per_rank() == [[], ["f0", "f0b"]] and sh["table_wise"].features_per_rank() == [0, 2]
    assert sh["row_wise"].feature_names_per_rank() == [["f1"], ["f1"]]
    seen = []
    sebc.register_post_lookup_tracker_fn(lambda feats, emb, module, extra: seen.append((feats.keys(), int(feats.values().numel()), emb, module)))
    sebc.register_post_odist_tracker_fn(lambda *a: None)
    kjt = KeyedJaggedTensor(keys=["f0", "f0b", "f1"], values=torch.tensor([1, 2, 3, 4, 5, 6]) + ctx.rank, lengths=torch.tensor([1, 1, 1, 1, 1, 1]))
    dmp(kjt)
    assert len(seen) == 1 and seen[0][2] is None and seen[0][3] is sebc
    if ctx.rank == 1:
        assert "f0" in seen[0][0] and seen[0][1] >= 8  # both ranks' f0 / f0b ids arrive at the owner of t0
    sebc.register_post_lookup_tracker_fn(lambda *a: seen.append("second"))  # replaces the first callback
    dmp(kjt)
    assert seen[-1] == "second" and len(seen) == 2


def test_sharded_embedding_module_interface():
    run_multi_process(_run_embedding_module_interface, world_size=2, backend="gloo")
