"""The reference-shaped composable sharding API (``parallel/sharding/*``, ``embedding_sharding.py``, ``embedding_lookup.py``, the per-group kernels):
every type's input dist -> lookup -> output dist reproduces the unsharded EmbeddingBagCollection / EmbeddingCollection, forward and fused backward."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _tables():
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, PoolingType

    return [
        EmbeddingBagConfig(name="t0", num_embeddings=37, embedding_dim=8, feature_names=["f0"]),
        EmbeddingBagConfig(name="t1", num_embeddings=50, embedding_dim=16, feature_names=["f1", "f2"]),
        EmbeddingBagConfig(name="t2", num_embeddings=23, embedding_dim=8, feature_names=["f3"], pooling=PoolingType.SUM),
    ]


def _batch(rank: int, keys, hash_sizes, B: int = 3, seed: int = 7):
    from torchrec_b200.sparse import KeyedJaggedTensor

    g = torch.Generator().manual_seed(seed + rank)
    lengths = torch.randint(0, 4, (len(keys) * B,), generator=g)
    vals = []
    for f, h in enumerate(hash_sizes):
        n = int(lengths[f * B : (f + 1) * B].sum())
        vals.append(torch.randint(0, h, (n,), generator=g))
    return KeyedJaggedTensor.from_lengths_sync(keys, torch.cat(vals), lengths)


def _infos(ebc, plan, tables, fused_params):
    from torchrec_b200.parallel.embedding_sharding import EmbeddingShardingInfo
    from torchrec_b200.parallel.embedding_types import EmbeddingTableConfig

    out = []
    for cfg in tables:
        tc = EmbeddingTableConfig(num_embeddings=cfg.num_embeddings, embedding_dim=cfg.embedding_dim, name=cfg.name, data_type=cfg.data_type,
                                  feature_names=cfg.feature_names, pooling=getattr(cfg, "pooling", None) or EmbeddingTableConfig.__dataclass_fields__["pooling"].default,
                                  embedding_names=list(cfg.feature_names))
        bag = ebc.embedding_bags if hasattr(ebc, "embedding_bags") else ebc.embeddings
        out.append(EmbeddingShardingInfo(tc, plan[cfg.name], bag[cfg.name].weight, fused_params))
    return out


def _pooled_types(ctx):
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embedding_sharding import EmbeddingShardingContext, KJTListSplitsAwaitable
    from torchrec_b200.parallel.sharding.cw_sharding import CwPooledEmbeddingSharding
    from torchrec_b200.parallel.sharding.dp_sharding import DpPooledEmbeddingSharding
    from torchrec_b200.parallel.sharding.grid_sharding import GridPooledEmbeddingSharding
    from torchrec_b200.parallel.sharding.rw_sharding import RwPooledEmbeddingSharding
    from torchrec_b200.parallel.sharding.tw_sharding import TwPooledEmbeddingSharding
    from torchrec_b200.parallel.sharding.twcw_sharding import TwCwPooledEmbeddingSharding
    from torchrec_b200.parallel.sharding.twrw_sharding import TwRwPooledEmbeddingSharding
    from torchrec_b200.parallel.types import ShardingEnv
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, rank, dev = ctx.world_size, ctx.rank, ctx.device
    L = ctx.local_size or W
    env = ShardingEnv.from_process_group(dist.group.WORLD)
    tables = _tables()
    keys = [f for t in tables for f in t.feature_names]
    hs = [t.num_embeddings for t in tables for _ in t.feature_names]
    lr = 0.5
    fused = {"optimizer": "exact_sgd", "learning_rate": lr}

    cases = {
        "tw": (TwPooledEmbeddingSharding, {"t0": sp.table_wise(rank=0), "t1": sp.table_wise(rank=W - 1), "t2": sp.table_wise(rank=1 % W)}, {}),
        "cw": (CwPooledEmbeddingSharding, {"t0": sp.column_wise(ranks=[W - 1, 0]), "t1": sp.column_wise(ranks=[0, 1 % W]), "t2": sp.column_wise(ranks=[1 % W])}, {"permute_embeddings": True}),
        "twcw": (TwCwPooledEmbeddingSharding, {"t0": sp.column_wise(ranks=[0, 0]), "t1": sp.column_wise(ranks=[1 % W, 0]), "t2": sp.column_wise(ranks=[0])}, {"permute_embeddings": True}),
        "rw": (RwPooledEmbeddingSharding, {"t0": sp.row_wise(), "t1": sp.row_wise(), "t2": sp.row_wise()}, {}),
        "rw_uneven": (RwPooledEmbeddingSharding, {"t0": sp.row_wise(([30] + [0] * (W - 2) + [7], "cpu")), "t1": sp.row_wise(), "t2": sp.row_wise()}, {}),
        "dp": (DpPooledEmbeddingSharding, {"t0": sp.data_parallel(), "t1": sp.data_parallel(), "t2": sp.data_parallel()}, {}),
    }
    if W // L >= 2:
        cases["twrw"] = (TwRwPooledEmbeddingSharding, {"t0": sp.table_row_wise(host_index=0), "t1": sp.table_row_wise(host_index=1), "t2": sp.table_row_wise(host_index=1)}, {})
        cases["grid"] = (GridPooledEmbeddingSharding, {"t0": sp.grid_shard(host_indexes=[1, 0]), "t1": sp.grid_shard(host_indexes=[0, 1]), "t2": sp.grid_shard(host_indexes=[0])}, {})

    for name, (cls, per_table, kw) in cases.items():
        torch.manual_seed(0)
        ebc = EmbeddingBagCollection(tables=tables, device=dev)
        golden = {t.name: ebc.embedding_bags[t.name].weight.detach().clone() for t in tables}
        plan = sp.construct_module_sharding_plan(ebc, per_table, local_size=L, world_size=W, device_type="cpu")
        sharding = cls(_infos(ebc, plan, tables, None if name == "dp" else fused), env, dev, **kw)
        in_dist, lookup, out_dist = sharding.create_input_dist(dev), sharding.create_lookup(dev), sharding.create_output_dist(dev)

        local = _batch(rank, keys, hs)
        feat_order = [keys.index(f) for f in sharding.feature_names()]
        routed = local.permute(feat_order)
        sctx = EmbeddingShardingContext()

        class Ctx:
            sharding_contexts = [sctx]

        dist_kjts = KJTListSplitsAwaitable([in_dist(routed)], Ctx()).wait().wait()
        pooled = out_dist(lookup(dist_kjts[0]), sctx).wait()

        # golden: the unsharded module on the same local batch; columns in the sharding's embedding_names() order
        kt = ebc(local)
        want = torch.cat([kt[n] for n in sharding.embedding_names()], dim=1)
        assert sharding.embedding_dims() == [kt[n].shape[1] for n in sharding.embedding_names()], name
        torch.testing.assert_close(pooled, want, rtol=1e-5, atol=1e-5, msg=lambda m: f"{name} forward: {m}")

        # backward: d(sum of all ranks' outputs) - every looked-up row moves by -lr * (number of times it was looked up, globally)
        if name == "dp":
            pooled.sum().backward()
            (pname, p), = list(lookup.named_parameters())
            assert p.grad is not None and pname == "t0_t1_t2.weight"
            continue
        pooled.sum().backward()
        all_vals = [None] * W
        dist.all_gather_object(all_vals, (local.values().tolist(), local.lengths().tolist()))
        B = local.stride()
        expect = {t.name: golden[t.name].clone() for t in tables}
        fi_table = [t.name for t in tables for _ in t.feature_names]
        for vals, lens in all_vals:
            o = 0
            for fi in range(len(keys)):
                n = sum(lens[fi * B : (fi + 1) * B])
                for v in vals[o : o + n]:
                    expect[fi_table[fi]][v] -= lr / W  # the output collectives divide gradients by the world size (comm_ops.set_gradient_division)
                o += n
        sd = lookup.state_dict()
        for g in sharding._grouped_embedding_configs:
            for t in g.embedding_tables:
                md = t.local_metadata
                r0, c0 = md.shard_offsets
                got = sd[f"{t.name}.weight"]
                if hasattr(got, "local_shards"):
                    got = next(s.tensor for s in got.local_shards() if list(s.metadata.shard_offsets) == [r0, c0])
                torch.testing.assert_close(got, expect[t.name][r0 : r0 + t.local_rows, c0 : c0 + t.local_cols], rtol=1e-5, atol=1e-5,
                                           msg=lambda m: f"{name} backward {t.name}@{r0},{c0}: {m}")
        assert [type(o).__name__ for o in lookup.fused_optimizers()] == ["EmbeddingFusedOptimizer"] * len(sharding._grouped_embedding_configs)


def _sequence_types(ctx):
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingCollection
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embedding import EmbeddingCollectionSharder
    from torchrec_b200.parallel.embedding_sharding import KJTListSplitsAwaitable
    from torchrec_b200.parallel.sharding.dp_sequence_sharding import DpSequenceEmbeddingSharding
    from torchrec_b200.parallel.sharding.rw_sequence_sharding import RwSequenceEmbeddingSharding
    from torchrec_b200.parallel.sharding.sequence_sharding import SequenceShardingContext
    from torchrec_b200.parallel.sharding.tw_sequence_sharding import TwSequenceEmbeddingSharding
    from torchrec_b200.parallel.types import ShardingEnv

    W, rank, dev = ctx.world_size, ctx.rank, ctx.device
    env = ShardingEnv.from_process_group(dist.group.WORLD)
    tables = [EmbeddingConfig(name="s0", num_embeddings=31, embedding_dim=8, feature_names=["f0", "f1"]),
              EmbeddingConfig(name="s1", num_embeddings=44, embedding_dim=8, feature_names=["f2"])]
    keys = ["f0", "f1", "f2"]
    hs = [31, 31, 44]
    cases = {
        "tw": (TwSequenceEmbeddingSharding, {"s0": sp.table_wise(rank=W - 1), "s1": sp.table_wise(rank=0)}),
        "rw": (RwSequenceEmbeddingSharding, {"s0": sp.row_wise(), "s1": sp.row_wise()}),
        "dp": (DpSequenceEmbeddingSharding, {"s0": sp.data_parallel(), "s1": sp.data_parallel()}),
    }
    for name, (cls, per_table) in cases.items():
        torch.manual_seed(0)
        ec = EmbeddingCollection(tables=tables, device=dev)
        plan = sp.construct_module_sharding_plan(ec, per_table, sharder=EmbeddingCollectionSharder(), local_size=W, world_size=W, device_type="cpu")
        sharding = cls(_infos(ec, plan, tables, None if name == "dp" else {"optimizer": "exact_sgd", "learning_rate": 0.1}), env, dev)
        in_dist, lookup, out_dist = sharding.create_input_dist(dev), sharding.create_lookup(dev), sharding.create_output_dist(dev)
        local = _batch(rank, keys, hs, B=4, seed=11)
        routed = local.permute([keys.index(f) for f in sharding.feature_names()])
        sctx = SequenceShardingContext(features_before_input_dist=routed)

        class Ctx:
            sharding_contexts = [sctx]

        dist_kjt = KJTListSplitsAwaitable([in_dist(routed)], Ctx()).wait().wait()[0]
        if name == "rw":
            sctx.unbucketize_permute_tensor = in_dist.unbucketize_permute_tensor
        if name == "dp":
            sctx.lengths_after_input_dist = dist_kjt.lengths()
        rows = out_dist(lookup(dist_kjt), sctx).wait()
        # golden rows in the routed feature order
        jts = ec(local)
        want = torch.cat([jts[f].values() for f in sharding.feature_names()], dim=0)
        torch.testing.assert_close(rows, want, rtol=1e-5, atol=1e-5, msg=lambda m: f"sequence {name}: {m}")
        rows.sum().backward()


def _write_dist_and_fused_splits(ctx):
    """Embedding update path (write dist + GroupedEmbeddingsUpdate) and the fused size exchange of several modules."""
    import torch.distributed as dist

    from torchrec_b200.modules.embedding_configs import EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingCollection
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embedding import EmbeddingCollectionSharder
    from torchrec_b200.parallel.embedding_sharding import FusedKJTListSplitsAwaitable, KJTListSplitsAwaitable, kjt_splits_meta
    from torchrec_b200.parallel.sharding.rw_sequence_sharding import RwSequenceEmbeddingSharding
    from torchrec_b200.parallel.sharding.sequence_sharding import SequenceShardingContext
    from torchrec_b200.parallel.types import ShardingEnv
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, rank, dev = ctx.world_size, ctx.rank, ctx.device
    env = ShardingEnv.from_process_group(dist.group.WORLD)
    tables = [EmbeddingConfig(name="u0", num_embeddings=20, embedding_dim=4, feature_names=["a", "b"])]
    torch.manual_seed(0)
    ec = EmbeddingCollection(tables=tables, device=dev)
    plan = sp.construct_module_sharding_plan(ec, {"u0": sp.row_wise()}, sharder=EmbeddingCollectionSharder(), local_size=W, world_size=W, device_type="cpu")
    sharding = RwSequenceEmbeddingSharding(_infos(ec, plan, tables, {"optimizer": "exact_sgd", "learning_rate": 0.1}), env, dev)
    update = sharding.create_update(dev)
    write = sharding.create_write_dist(dev)
    # rank r rewrites ids {r, 10 + r} of feature a and {5 + r} of feature b with recognisable rows
    ids = KeyedJaggedTensor.from_lengths_sync(["a", "b"], torch.tensor([rank, 10 + rank, 5 + rank]), torch.tensor([2, 0, 0, 1]))
    rows = torch.stack([torch.full((4,), 100.0 + rank), torch.full((4,), 200.0 + rank), torch.full((4,), 300.0 + rank)])
    update(write(ids, rows).wait().wait())
    sd = update._lookup.state_dict()["u0.weight"]
    local = sd.local_shards()[0] if hasattr(sd, "local_shards") else None
    w = local.tensor if local is not None else sd
    r0 = local.metadata.shard_offsets[0] if local is not None else 0
    block = -(-20 // W)
    for r in range(W):
        for gid, val in ((r, 100.0 + r), (10 + r, 200.0 + r), (5 + r, 300.0 + r)):
            if gid // block == rank:
                assert float(w[gid - r0, 0]) == val, (gid, float(w[gid - r0, 0]), val)

    # fused splits: two "modules" exchange their sizes in ONE collective and still receive the right KJTs
    a = KeyedJaggedTensor.from_lengths_sync(["x"] * 0 + [f"k{i}" for i in range(W)], torch.arange(2 * W) + 100 * rank, torch.ones(2 * W, dtype=torch.int64))
    b = KeyedJaggedTensor.from_lengths_sync([f"m{i}" for i in range(W)], torch.arange(W) + 1000 * rank, torch.ones(W, dtype=torch.int64))

    class C1:
        sharding_contexts = [SequenceShardingContext()]

    class C2:
        sharding_contexts = [SequenceShardingContext()]

    reqs = [KJTListSplitsAwaitable([kjt_splits_meta(dist.group.WORLD, a, [1] * W)], C1()), KJTListSplitsAwaitable([kjt_splits_meta(dist.group.WORLD, b, [1] * W)], C2())]
    got = [aw.wait() for aw in FusedKJTListSplitsAwaitable(reqs, [C1(), C2()], dist.group.WORLD).wait()]
    ka, kb = got[0][0], got[1][0]
    assert ka.keys() == [f"k{rank}"] and kb.keys() == [f"m{rank}"]
    assert ka.values().tolist() == [v for r in range(W) for v in (2 * rank + 100 * r, 2 * rank + 1 + 100 * r)]
    assert kb.values().tolist() == [rank + 1000 * r for r in range(W)]


def test_pooled_shardings_2_ranks():
    run_multi_process(_pooled_types, world_size=2)


def test_pooled_shardings_two_hosts_of_two():
    run_multi_process(_pooled_types, world_size=4, local_size=2)


def test_sequence_shardings_2_ranks():
    run_multi_process(_sequence_types, world_size=2)


def test_write_dist_and_fused_splits():
    run_multi_process(_write_dist_and_fused_splits, world_size=2)


def test_group_tables_and_kernels_single_process():
    from torchrec_b200.modules.embedding_configs import DataType, PoolingType
    from torchrec_b200.parallel.embedding_sharding import bucketize_kjt_before_all2all, group_tables
    from torchrec_b200.parallel.embedding_types import EmbeddingComputeKernel as K
    from torchrec_b200.parallel.embedding_types import ShardedEmbeddingTable
    from torchrec_b200.parallel.quant_embedding_kernel import QuantBatchedEmbeddingBag
    from torchrec_b200.parallel.batched_embedding_kernel import BatchedFusedEmbeddingBag
    from torchrec_b200.sparse import KeyedJaggedTensor

    def tb(name, dim, k, pooling=PoolingType.SUM, dt=DataType.FP32, fp=None):
        return ShardedEmbeddingTable(num_embeddings=16, embedding_dim=dim, name=name, feature_names=[name + "_f"], embedding_names=[name + "_f"], data_type=dt,
                                     pooling=pooling, compute_kernel=k, local_rows=16, local_cols=dim, fused_params=fp)

    groups = group_tables([[tb("a", 8, K.FUSED), tb("b", 16, K.FUSED), tb("c", 8, K.FUSED, PoolingType.MEAN), tb("d", 8, K.DENSE), tb("e", 8, K.FUSED, dt=DataType.FP16),
                            tb("f", 8, K.FUSED_UVM_CACHING, fp={"cache_load_factor": 0.5}), tb("g", 8, K.FUSED, fp={"cache_load_factor": 0.1})], []])
    assert len(groups) == 2 and groups[1] == []
    names = [g.table_names() for g in groups[0]]
    assert names == [["a", "b", "f", "g"], ["c"], ["d"], ["e"]], names
    assert groups[0][0].compute_kernel == K.FUSED_UVM_CACHING and abs(groups[0][0].fused_params["cache_load_factor"] - 0.3) < 1e-9

    kjt = KeyedJaggedTensor.from_lengths_sync(["x", "y"], torch.tensor([0, 9, 15, 3, 12]), torch.tensor([2, 1, 1, 1]))
    out, unb = bucketize_kjt_before_all2all(kjt, 2, torch.tensor([8, 8]), output_permute=True)
    assert out.keys() == ["x", "y", "x", "y"] and out.lengths().tolist() == [1, 0, 1, 0, 1, 1, 0, 1]
    assert out.values().tolist() == [0, 3, 1, 7, 4] and unb.tolist() == [0, 2, 3, 1, 4]

    # float kernel -> quantized kernel
    g = groups[0][1]
    g.fused_params = {"optimizer": "exact_sgd", "learning_rate": 0.1}
    fk = BatchedFusedEmbeddingBag(g, device=torch.device("cpu"))
    q = QuantBatchedEmbeddingBag.from_float(fk)
    f = KeyedJaggedTensor.from_lengths_sync(["c_f"], torch.tensor([1, 2, 3]), torch.tensor([2, 1]))
    torch.testing.assert_close(q(f), fk(f).detach(), rtol=0.05, atol=0.02)
