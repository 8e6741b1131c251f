"""Row-wise sharded TensorPool / KeyedJaggedTensorPool on 2 gloo ranks (reference tests/test_tensor_pool.py,
test_keyed_jagged_tensor_pool.py)."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _run(ctx, replicated: bool):
    from torchrec_b200.modules.object_pool import KeyedJaggedTensorPool, TensorPool
    from torchrec_b200.parallel.object_pool import KeyedJaggedTensorPoolSharder, ObjectPoolShardingPlan, ObjectPoolShardingType, TensorPoolSharder
    from torchrec_b200.parallel.types import ShardingEnv
    from torchrec_b200.sparse import KeyedJaggedTensor
    import torch.distributed as dist

    env = ShardingEnv.from_process_group(dist.group.WORLD)
    st = ObjectPoolShardingType.REPLICATED_ROW_WISE if replicated else ObjectPoolShardingType.ROW_WISE
    tp = TensorPool(11, 3, torch.float32)
    sp_ = TensorPoolSharder().shard(tp, ObjectPoolShardingPlan(st), env, ctx.device)
    # every rank writes its own ids; afterwards everybody can read everything
    ids = torch.tensor([0, 5, 9]) if ctx.rank == 0 else torch.tensor([10, 2, 6])
    sp_.update(ids, ids.float().unsqueeze(1).repeat(1, 3) + 0.5)
    q = torch.tensor([9, 10, 0, 2, 3, 6, 5])
    out = sp_.lookup(q)
    exp = torch.where(torch.tensor([True, True, True, True, False, True, True]).unsqueeze(1), q.float().unsqueeze(1).repeat(1, 3) + 0.5, torch.zeros(7, 3))
    torch.testing.assert_close(out, exp)
    if not replicated:
        kp = KeyedJaggedTensorPool(8, {"a": 3, "b": 2})
        skp = KeyedJaggedTensorPoolSharder().shard(kp, ObjectPoolShardingPlan(st), env, ctx.device)
        if ctx.rank == 0:
            skp.update(torch.tensor([1, 6]), KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([11, 12, 61, 13]), lengths=torch.tensor([2, 1, 1, 0])))
        else:
            skp.update(torch.tensor([4]), KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([41, 42, 43, 44]), lengths=torch.tensor([3, 1])))
        got = skp.lookup(torch.tensor([6, 4, 1, 0])).to_dict()
        assert got["a"].to_dense()[0].tolist() == [61] and got["a"].to_dense()[1].tolist() == [41, 42, 43] and got["a"].to_dense()[2].tolist() == [11, 12]
        assert got["b"].lengths().tolist() == [0, 1, 1, 0] and got["b"].values().tolist() == [44, 13]


@pytest.mark.parametrize("replicated", [False, True])
def test_sharded_pools(replicated):
    run_multi_process(_run, world_size=2, backend="gloo", replicated=replicated)


def _composable_pool_shardings(ctx):
    """The decomposed pool shardings (ids dist + values dists + local lookups) reproduce a replicated reference pool."""
    import torch
    import torch.distributed as dist

    from torchrec_b200.modules.object_pool_lookups import TensorJaggedIndexSelectLookup, TensorLookup
    from torchrec_b200.parallel.sharding.rw_kjt_pool_sharding import KeyedJaggedTensorPoolRwSharding
    from torchrec_b200.parallel.sharding.rw_tensor_pool_sharding import TensorPoolRwSharding
    from torchrec_b200.parallel.types import ShardingEnv
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, rank, dev = ctx.world_size, ctx.rank, ctx.device
    env = ShardingEnv.from_process_group(dist.group.WORLD)
    pool_size, dim = 11, 4
    sh = TensorPoolRwSharding(pool_size, dim, env, dev)
    lookup = TensorLookup(sh.local_pool_size, dim, torch.float32, dev)
    # update: rank r writes ids {r, 5 + r, 10 - r}
    ids = torch.tensor([rank, 5 + rank, 10 - rank])
    vals = torch.stack([torch.full((dim,), float(100 * rank + i)) for i in ids.tolist()])
    c = sh.create_context()
    local_ids = sh.create_update_ids_dist()(c, ids).wait().wait()
    lookup.update(local_ids, sh.create_update_values_dist()(c, vals).wait())
    dist.barrier()
    q = torch.tensor([10 - rank, rank, 5 + rank, rank])
    c2 = sh.create_context()
    got = sh.create_lookup_values_dist()(c2, lookup.lookup(sh.create_lookup_ids_dist()(c2, q).wait().wait())).wait()
    want = torch.stack([torch.full((dim,), float(100 * rank + i)) for i in q.tolist()])
    torch.testing.assert_close(got, want)

    fml = {"a": 3, "b": 2}
    ks = KeyedJaggedTensorPoolRwSharding(env, dev, pool_size, 2, fml)
    klookup = TensorJaggedIndexSelectLookup(ks.local_pool_size, torch.int64, fml, False, dev)
    kids = torch.tensor([rank, 6 + rank])
    kjt = KeyedJaggedTensor.from_lengths_sync(["a", "b"], torch.tensor([10 + rank, 11 + rank, 20 + rank, 30 + rank, 40 + rank, 41 + rank]), torch.tensor([2, 1, 1, 2]))
    kc = ks.create_context()
    lids = ks.create_update_ids_dist()(kc, kids).wait().wait()
    klookup.update(lids, ks.create_update_values_dist()(kc, kjt).wait())
    dist.barrier()
    other = (rank + 1) % W
    kq = torch.tensor([6 + other, rank])
    kc2 = ks.create_context()
    out = ks.create_lookup_values_dist()(kc2, klookup.lookup(ks.create_lookup_ids_dist()(kc2, kq).wait().wait())).wait()
    assert out.keys() == ["a", "b"] and out.lengths().tolist() == [1, 2, 2, 1]
    assert out.values().tolist() == [20 + other, 10 + rank, 11 + rank, 40 + other, 41 + other, 30 + rank]


def test_composable_pool_shardings():
    run_multi_process(_composable_pool_shardings, world_size=2)


def test_inference_pools_over_local_devices():
    """Serving form of the pools: one process, the rows block-partitioned over W local devices (cpu stand-ins here); lookups return the
    unsharded pool's answer in the caller's order; updates are refused."""
    from torchrec_b200.modules.object_pool import KeyedJaggedTensorPool, TensorPool
    from torchrec_b200.parallel.keyed_jagged_tensor_pool import ShardedInferenceKeyedJaggedTensorPool
    from torchrec_b200.parallel.object_pool import KeyedJaggedTensorPoolSharder, ObjectPoolShardingPlan, ObjectPoolShardingType, TensorPoolSharder
    from torchrec_b200.parallel.tensor_pool import LocalShardPool, ShardedInferenceTensorPool
    from torchrec_b200.parallel.types import ShardingEnv
    from torchrec_b200.sparse import KeyedJaggedTensor

    env = ShardingEnv.from_local(world_size=3, rank=0)
    plan = ObjectPoolShardingPlan(ObjectPoolShardingType.ROW_WISE, inference=True)
    tp = TensorPool(11, 4, torch.float32)
    tp.update(torch.arange(11), torch.arange(44, dtype=torch.float32).view(11, 4))
    stp = TensorPoolSharder().shard(tp, plan, env, torch.device("cpu"))
    assert isinstance(stp, ShardedInferenceTensorPool) and [p._shard.shape[0] for p in stp._local_shard_pools] == [4, 4, 3] and isinstance(stp._local_shard_pools[0], LocalShardPool)
    q = torch.tensor([10, 0, 7, 3, 4, 10, 8])
    torch.testing.assert_close(stp(q), tp.lookup(q))
    assert stp.lookup(torch.zeros(0, dtype=torch.long)).shape == (0, 4) and (stp.pool_size, stp.dim) == (11, 4)
    with pytest.raises(NotImplementedError):
        stp.update(q, torch.zeros(7, 4))
    kp = KeyedJaggedTensorPool(8, {"a": 3, "b": 2})
    kp.update(torch.tensor([1, 6, 4]), KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([11, 12, 61, 41, 42, 43, 13, 44]), lengths=torch.tensor([2, 1, 3, 1, 0, 1])))
    skp = KeyedJaggedTensorPoolSharder().shard(kp, plan, env, torch.device("cpu"))
    assert isinstance(skp, ShardedInferenceKeyedJaggedTensorPool)
    ids = torch.tensor([6, 4, 1, 0, 7])
    got, want = skp(ids), kp.lookup(ids)
    assert got.keys() == want.keys() and torch.equal(got.values(), want.values()) and torch.equal(got.lengths(), want.lengths())
    with pytest.raises(NotImplementedError):
        skp.update(ids, got)
    # the training form is still what a plan without the inference flag gives
    assert type(TensorPoolSharder().shard(tp, ObjectPoolShardingPlan(ObjectPoolShardingType.ROW_WISE), ShardingEnv.from_local(1, 0), torch.device("cpu"))).__name__ == "ShardedTensorPool"
