"""UVM-cached table-batched embeddings: training through a small HBM cache must equal training the full table."""
import pytest
import torch


def _run(device, algo):
    from torchrec_b200.ops.tbe import OptimType, PoolingMode, TableBatchedEmbeddingBags
    from torchrec_b200.ops.uvm import UvmCachedEmbeddingBags

    torch.manual_seed(0)
    specs, fmap = [(300, 8), (200, 16)], [0, 1, 1]
    kw = dict(pooling_mode=PoolingMode.SUM, optimizer=OptimType.EXACT_ROWWISE_ADAGRAD, learning_rate=0.1, eps=1e-3)
    full = TableBatchedEmbeddingBags(specs, fmap, device=device, **kw)
    cached = UvmCachedEmbeddingBags(specs, fmap, cache_load_factor=0.1, cache_algorithm=algo, min_cache_rows=40, device=device, **kw)
    cached.init_parameters()
    with torch.no_grad():
        for w_c, w_f in zip(cached.split_embedding_weights(), full.split_embedding_weights()):
            w_c.copy_(w_f.cpu())
    cached.load_rows_changed()
    assert cached.cache_rows == [40, 40]
    g = torch.Generator().manual_seed(1)
    B = 6
    for step in range(25):
        lengths = torch.randint(0, 3, (3 * B,), generator=g)
        off = torch.cat([lengths.new_zeros(1), lengths.cumsum(0)]).to(device)
        rows = torch.tensor([300, 200, 200]).repeat_interleave(B).repeat_interleave(lengths)
        idx = (torch.rand(int(lengths.sum()), generator=g) * rows).long().to(device)
        if step % 2 == 0:
            cached.prefetch(idx, off, B)  # explicit prefetch (pipeline) and implicit (inside forward) both work
        o_f = full(idx, off, None, batch_size=B)
        o_c = cached(idx, off, None, batch_size=B)
        torch.testing.assert_close(o_c, o_f)
        gout = torch.randn(o_f.shape, generator=g).to(device)
        o_f.backward(gout)
        o_c.backward(gout)
    assert cached.stats["evictions"] > 0 and cached.stats["hits"] > 0
    for w_c, w_f in zip(cached.split_embedding_weights(), full.split_embedding_weights()):
        torch.testing.assert_close(w_c, w_f.cpu())
    for s_c, s_f in zip(cached.split_optimizer_states(), full.split_optimizer_states()):
        torch.testing.assert_close(s_c["momentum1"], s_f["momentum1"].cpu())


@pytest.mark.parametrize("algo", ["lru", "lfu"])
def test_uvm_cache_cpu(algo):
    _run(torch.device("cpu"), algo)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["lru", "lfu"])
def test_uvm_cache_gpu(algo):
    _run(torch.device("cuda:0"), algo)


@pytest.mark.gpu
def test_managed_zero_copy_location_gpu():
    """Tables in pinned host memory read / updated zero-copy by the same kernels."""
    from torchrec_b200.ops.tbe import EmbeddingLocation, OptimType, PoolingMode, TableBatchedEmbeddingBags

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    kw = dict(pooling_mode=PoolingMode.SUM, optimizer=OptimType.EXACT_ROWWISE_ADAGRAD, learning_rate=0.1, eps=1e-3, device=dev)
    a = TableBatchedEmbeddingBags([(500, 32)], [0], **kw)
    b = TableBatchedEmbeddingBags([(500, 32)], [0], location=EmbeddingLocation.MANAGED, **kw)
    assert not b.weights.is_cuda and b.weights.is_pinned()
    b.weights.copy_(a.weights.cpu())
    idx = torch.randint(0, 500, (64,), device=dev)
    off = torch.arange(0, 65, 2, device=dev)
    for _ in range(3):
        oa, ob = a(idx, off, None, batch_size=32), b(idx, off, None, batch_size=32)
        torch.testing.assert_close(oa, ob)
        g = torch.randn_like(oa)
        oa.backward(g)
        ob.backward(g)
    torch.cuda.synchronize()
    torch.testing.assert_close(b.weights, a.weights.cpu())
    b.to(dev)
    assert not b.weights.is_cuda  # .to() leaves zero-copy tables on the host


def test_sharded_ebc_with_uvm_caching_kernel(monkeypatch):
    """compute_kernel=fused_uvm_caching in the plan -> cached tables behind the sharded EBC (CPU emulation of the cache)."""
    monkeypatch.setenv("TRB_UVM_ON_CPU", "1")
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.ops.uvm import UvmCachedEmbeddingBags
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import CacheParams, ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    torch.manual_seed(0)
    tables = [EmbeddingBagConfig(name="big", embedding_dim=8, num_embeddings=5000, feature_names=["f0"]),
              EmbeddingBagConfig(name="small", embedding_dim=8, num_embeddings=50, feature_names=["f1"])]
    gold = EmbeddingBagCollection(tables)
    local = EmbeddingBagCollection(tables)
    local.load_state_dict(gold.state_dict())
    apply_optimizer_in_backward(torch.optim.SGD, local.parameters(), {"lr": 0.5})
    plan = sp.construct_module_sharding_plan(local, {"big": sp.table_wise(rank=0, compute_kernel="fused_uvm_caching"), "small": sp.table_wise(rank=0)},
                                             sharder=EmbeddingBagCollectionSharder(), world_size=1, local_size=1, device_type="cpu")
    plan["big"].cache_params = CacheParams(load_factor=0.05)

    class W(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = local

        def forward(self, k):
            return self.ebc(k).values()

    m = DistributedModelParallel(W(), device=torch.device("cpu"), plan=ShardingPlan({"ebc": plan}), sharders=[EmbeddingBagCollectionSharder()])
    eng = m.module.ebc.engine
    assert any(isinstance(t, UvmCachedEmbeddingBags) for t in eng._tbes)
    opt = torch.optim.SGD(gold.parameters(), lr=0.5)
    g = torch.Generator().manual_seed(3)
    for _ in range(6):
        kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=torch.cat([torch.randint(0, 5000, (8,), generator=g), torch.randint(0, 50, (8,), generator=g)]), lengths=torch.full((8,), 2))
        out, ref = m(kjt), gold(kjt).values()
        torch.testing.assert_close(out, ref)
        out.sum().backward()
        opt.zero_grad()
        ref.sum().backward()
        opt.step()
    sd = m.state_dict()
    w = sd["ebc.embedding_bags.big.weight"]
    w = w.local_shards()[0].tensor if hasattr(w, "local_shards") else w
    torch.testing.assert_close(w, gold.embedding_bags["big"].weight.detach())
