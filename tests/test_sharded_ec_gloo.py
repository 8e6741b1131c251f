"""ShardedEmbeddingCollection (sequence embeddings) vs the unsharded module on 2 gloo ranks."""
import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _run(ctx, sharding: str, dedup: bool, grad_div: bool = False):
    from torchrec_b200.modules.embedding_configs import EmbeddingConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.parallel.embedding import EmbeddingCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    set_gradient_division(grad_div)
    torch.manual_seed(0)
    W, B = ctx.world_size, 5
    tables = lambda: [EmbeddingConfig(name="t0", embedding_dim=8, num_embeddings=30, feature_names=["f0", "f1"]),
                      EmbeddingConfig(name="t1", embedding_dim=8, num_embeddings=40, feature_names=["f2"])]
    gold = EmbeddingCollection(tables())
    local = EmbeddingCollection(tables())
    local.load_state_dict(gold.state_dict())
    apply_optimizer_in_backward(torch.optim.SGD, local.parameters(), {"lr": 0.1})
    gens = {"tw": {"t0": sp.table_wise(rank=0), "t1": sp.table_wise(rank=1)},
            "rw": {"t0": sp.row_wise(), "t1": sp.row_wise()},
            "cw": {"t0": sp.column_wise(ranks=[0, 1]), "t1": sp.column_wise(ranks=[1, 0])},
            "mixed": {"t0": sp.row_wise(), "t1": sp.table_wise(rank=1)}}[sharding]
    sharder = EmbeddingCollectionSharder(use_index_dedup=dedup)
    plan = sp.construct_module_sharding_plan(local, gens, sharder=sharder, world_size=W, local_size=W, device_type="cpu")

    class Wrap(torch.nn.Module):
        def __init__(self, ec):
            super().__init__()
            self.ec = ec

        def forward(self, kjt):
            return self.ec(kjt)

    model = DistributedModelParallel(Wrap(local), device=torch.device("cpu"), plan=ShardingPlan({"ec": plan}), sharders=[sharder])
    # gradient division: the sequence all-to-all divides by W in backward, i.e. the sharded update uses lr / W
    gold_opt = torch.optim.SGD(gold.parameters(), lr=0.1 / (ctx.world_size if grad_div else 1))

    def batch(seed):
        g = torch.Generator().manual_seed(seed)
        lengths = torch.randint(0, 4, (3 * B,), generator=g)
        hs = [30, 30, 40]
        vals = torch.cat([torch.randint(0, hs[i], (int(lengths[i * B : (i + 1) * B].sum()),), generator=g) for i in range(3)])
        return KeyedJaggedTensor(keys=["f0", "f1", "f2"], values=vals, lengths=lengths)

    for step in range(2):
        batches = [batch(100 * step + r) for r in range(W)]
        out = model(batches[ctx.rank])
        out = {k: out[k] for k in ["f0", "f1", "f2"]}
        gouts = [gold(b) for b in batches]
        loss = 0
        for k in ["f0", "f1", "f2"]:
            torch.testing.assert_close(out[k].values(), gouts[ctx.rank][k].values(), rtol=1e-5, atol=1e-6)
            assert torch.equal(out[k].lengths(), gouts[ctx.rank][k].lengths())
            w = torch.linspace(0.5, 1.5, 8)
            loss = loss + (out[k].values() * w).sum()
        loss.backward()
        gold_opt.zero_grad()
        sum((go[k].values() * torch.linspace(0.5, 1.5, 8)).sum() for go in gouts for k in go).backward()
        gold_opt.step()
    sd = model.state_dict()
    for name in ["t0", "t1"]:
        st = sd[f"ec.embeddings.{name}.weight"]
        ref = gold.embeddings[name].weight.detach()
        for sh in st.local_shards():
            o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
            torch.testing.assert_close(sh.tensor, ref[o[0] : o[0] + s[0], o[1] : o[1] + s[1]], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("sharding", ["tw", "rw", "cw", "mixed"])
def test_sharded_ec_matches_unsharded(sharding):
    run_multi_process(_run, world_size=2, backend="gloo", sharding=sharding, dedup=False)


def test_sharded_ec_with_index_dedup():
    run_multi_process(_run, world_size=2, backend="gloo", sharding="rw", dedup=True)


@pytest.mark.parametrize("sharding", ["tw", "rw"])
def test_sharded_ec_gradient_division(sharding):
    """Default setting (gradient division on): sequence embedding gradients are divided by the world size like the pooled path."""
    run_multi_process(_run, world_size=2, backend="gloo", sharding=sharding, dedup=False, grad_div=True)
