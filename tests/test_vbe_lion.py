"""Variable batch per feature (VBE) through the sharded EBC; LION fused optimizer reference semantics (CPU)."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def test_lion_matches_reference_formula():
    from torchrec_b200.ops.tbe import OptimType, PoolingMode, TableBatchedEmbeddingBags

    torch.manual_seed(0)
    t = TableBatchedEmbeddingBags([(10, 4)], [0], pooling_mode=PoolingMode.SUM, optimizer=OptimType.LION, learning_rate=0.1, beta1=0.9, beta2=0.99)
    t.init_parameters([(-1, 1)])
    w0 = t.split_embedding_weights()[0].clone()
    idx, off = torch.tensor([1, 3, 3]), torch.tensor([0, 1, 3])
    g_out = torch.tensor([[1.0, -2.0, 0.5, 0.0], [0.25, 0.5, -1.0, 2.0]])
    m = torch.zeros(10, 4)
    w = w0.clone()
    for _ in range(2):
        out = t(idx, off, None, batch_size=2)
        out.backward(g_out)
        g = torch.zeros(10, 4)
        g[1] += g_out[0]
        g[3] += 2 * g_out[1]
        touched = torch.tensor([1, 3])
        c = 0.9 * m[touched] + 0.1 * g[touched]
        w[touched] -= 0.1 * torch.sign(c)
        m[touched] = 0.99 * m[touched] + 0.01 * g[touched]
    torch.testing.assert_close(t.split_embedding_weights()[0], w)
    torch.testing.assert_close(t.split_optimizer_states()[0]["momentum1"], m)


def _run_vbe(ctx, sharding: str):
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    torch.manual_seed(0)
    W, dev = ctx.world_size, ctx.device
    tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=40, feature_names=["f0"]),
              EmbeddingBagConfig(name="t1", embedding_dim=4, num_embeddings=30, feature_names=["f1"])]
    gold = EmbeddingBagCollection(tables, device=dev)
    local = EmbeddingBagCollection(tables, device=dev)
    local.load_state_dict(gold.state_dict())
    apply_optimizer_in_backward(torch.optim.SGD, local.parameters(), {"lr": 0.1})
    gens = {"tw": {"t0": sp.table_wise(rank=0), "t1": sp.table_wise(rank=1)}, "rw": {"t0": sp.row_wise(), "t1": sp.row_wise()}}[sharding]
    plan = sp.construct_module_sharding_plan(local, gens, sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type="cpu")

    class Wrap(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = local

        def forward(self, k):
            return self.ebc(k).values()

    model = DistributedModelParallel(Wrap(), device=dev, plan=ShardingPlan({"ebc": plan}), sharders=[EmbeddingBagCollectionSharder()])
    g = torch.Generator().manual_seed(ctx.rank)
    # f0 has 2 distinct bags, f1 has 3; the full batch is 4 (rank 1: 5 to exercise uneven batches)
    Bf = 4 + ctx.rank
    l0, l1 = torch.randint(1, 3, (2,), generator=g), torch.randint(0, 3, (3,), generator=g)
    kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.cat([torch.randint(0, 40, (int(l0.sum()),), generator=g), torch.randint(0, 30, (int(l1.sum()),), generator=g)]),
                            lengths=torch.cat([l0, l1]), stride_per_key_per_rank=[[2], [3]],
                            inverse_indices=(["f0", "f1"], torch.stack([torch.randint(0, 2, (Bf,), generator=g), torch.randint(0, 3, (Bf,), generator=g)])))
    out = model(kjt)
    ref = gold(kjt).values()
    assert out.shape == (Bf, 12)
    torch.testing.assert_close(out, ref)
    out.sum().backward()  # gradient flows back through the expansion into the fused backward
    out2 = model(kjt)
    assert not torch.allclose(out2, ref)


@pytest.mark.parametrize("sharding", ["tw", "rw"])
def test_sharded_vbe(sharding):
    run_multi_process(_run_vbe, world_size=2, backend="gloo", sharding=sharding)
