"""Row-wise sharded TensorPool / KeyedJaggedTensorPool on 2 gloo ranks (reference tests/test_tensor_pool.py,
test_keyed_jagged_tensor_pool.py)."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _run(ctx, replicated: bool):
    from torchrec_b200.modules.object_pools import KeyedJaggedTensorPool, TensorPool
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
