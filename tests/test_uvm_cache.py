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
