"""Sharded checkpoint: save under one plan, load under another (re-sharding on load), 2 gloo ranks."""
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _run(ctx, tmp: str):
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import checkpoint as ckpt
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    W, dev = ctx.world_size, ctx.device
    tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=40, feature_names=["f0"]),
              EmbeddingBagConfig(name="t1", embedding_dim=16, num_embeddings=30, feature_names=["f1"])]

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = EmbeddingBagCollection(tables, device=dev)
            self.lin = torch.nn.Linear(24, 1)

        def forward(self, k):
            return self.lin(self.ebc(k).values())

    def build(gens, seed):
        torch.manual_seed(seed)
        m = M()
        apply_optimizer_in_backward(RowWiseAdagrad, m.ebc.parameters(), {"lr": 0.1})
        plan = sp.construct_module_sharding_plan(m.ebc, gens, sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type="cpu")
        return DistributedModelParallel(m, device=dev, plan=ShardingPlan({"ebc": plan}), sharders=[EmbeddingBagCollectionSharder()])

    a = build({"t0": sp.table_wise(rank=0), "t1": sp.row_wise()}, seed=0)
    g = torch.Generator().manual_seed(5 + ctx.rank)
    kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.cat([torch.randint(0, 40, (6,), generator=g), torch.randint(0, 30, (6,), generator=g)]), lengths=torch.full((12,), 1))
    for _ in range(2):
        a(kjt).sum().backward()
    ckpt.save(a, a.fused_optimizer, tmp, extra={"step": 2})
    c = build({"t0": sp.row_wise(), "t1": sp.column_wise(ranks=[1, 0])}, seed=7)  # row-wise state folds onto column shards
    ckpt.load(c, c.fused_optimizer, tmp)
    torch.testing.assert_close(c(kjt), a(kjt))
    b = build({"t0": sp.row_wise(), "t1": sp.table_wise(rank=1)}, seed=123)  # different plan, different init
    extra = ckpt.load(b, b.fused_optimizer, tmp)
    assert extra == {"step": 2}
    torch.testing.assert_close(b(kjt), a(kjt))
    torch.testing.assert_close(b.module.lin.weight, a.module.lin.weight)
    # optimizer state came along: one more identical step keeps the models identical
    a(kjt).sum().backward()
    b(kjt).sum().backward()
    torch.testing.assert_close(b(kjt), a(kjt), rtol=1e-5, atol=1e-6)


def test_checkpoint_reshard_on_load(tmp_path):
    run_multi_process(_run, world_size=2, backend="gloo", tmp=str(tmp_path))
