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


def test_stochastic_rounding_is_unbiased_cpu_reference():
    """bf16 / fp16 tables: an update far below half an ulp vanishes with round-to-nearest and survives in expectation with SR."""
    import torch

    from torchrec_b200.ops.tbe import OptimType, TableBatchedEmbeddingBags, stochastic_round

    x = torch.full((200000,), 1.0 + 2.0**-10)       # bf16 ulp at 1.0 is 2^-7: x sits 1/8 of the way to the next value
    r = stochastic_round(x, torch.bfloat16, torch.Generator().manual_seed(0)).float()
    assert set(r.unique().tolist()) == {1.0, 1.0 + 2.0**-7}
    assert abs(r.mean().item() - x[0].item()) < 2e-5
    xh = torch.full((200000,), 1.0 + 2.0**-12)       # fp16 ulp at 1.0 is 2^-10: 1/4 of the way
    rh = stochastic_round(xh, torch.float16, torch.Generator().manual_seed(1)).float()
    assert abs(rh.mean().item() - xh[0].item()) < 1e-5 and rh.unique().numel() == 2

    def run(sr: bool) -> float:
        torch.manual_seed(0)
        tbe = TableBatchedEmbeddingBags([(4, 64)], optimizer=OptimType.EXACT_SGD, learning_rate=1.0, weights_precision=torch.bfloat16, output_dtype=torch.float32,
                                        stochastic_rounding=sr, sr_seed=7)
        with torch.no_grad():
            tbe.weights.fill_(1.0)
        idx, off = torch.arange(4), torch.arange(5)
        for _ in range(64):
            tbe(idx, off).backward(torch.full((4, 64), -2.0**-11))   # each step wants +2^-11, 1/16 of a bf16 ulp
        return tbe.weights.float().mean().item()

    assert run(False) == 1.0                                  # round-to-nearest: every update is lost
    assert abs(run(True) - (1.0 + 64 * 2.0**-11)) < 6e-3      # SR: the mean moves by the requested 2^-5 (noise ~ 1e-3)


def test_vbe_kjt_permute_split_to_dict():
    """Variable-batch KJT: permute must move the per-key LENGTH runs (not the stride table) along with the values."""
    import torch

    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    kjt = KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([10, 20, 21, 30, 31, 32]), lengths=torch.tensor([1, 2, 3]),
                            stride_per_key_per_rank=[[2], [1]])
    p = kjt.permute([1, 0])
    assert p.keys() == ["b", "a"]
    assert p.lengths().tolist() == [3, 1, 2]
    assert p.values().tolist() == [30, 31, 32, 10, 20, 21]
    assert p.stride_per_key() == [1, 2]
    d = p.to_dict()
    assert d["a"].lengths().tolist() == [1, 2] and d["a"].values().tolist() == [10, 20, 21]
    assert d["b"].lengths().tolist() == [3] and d["b"].values().tolist() == [30, 31, 32]
    first, second = p.split([1, 1])
    assert first.keys() == ["b"] and first.lengths().tolist() == [3] and second.values().tolist() == [10, 20, 21]
    rep = kjt.permute([0, 1, 0])  # repeated keys (column-wise sharding replicates features)
    assert rep.lengths().tolist() == [1, 2, 3, 1, 2] and rep.values().tolist() == [10, 20, 21, 30, 31, 32, 10, 20, 21]
