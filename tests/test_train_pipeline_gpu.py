"""Train pipelines ON CUDA (real memcpy / data-dist / prefetch streams, pinned host batches, the NVLink sparse plane on one rank): every
pipeline must reproduce the plain training loop step by step - losses and final weights - which is what proves the stream / event /
record_stream discipline (methodology of the reference's train_pipeline/tests/test_train_pipelines.py). A second process group-free
variant drives uvm-cached tables through the prefetch pipeline."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(device, seed=0, kernel=None, dims=64):
    from torchrec_b200.models.dlrm import DLRM, DLRMTrain
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.keyed import CombinedOptimizer, KeyedOptimizerWrapper
    from torchrec_b200.optim.optimizers import in_backward_optimizer_filter
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    torch.manual_seed(seed)
    keys = [f"f{i}" for i in range(5)]
    hashes = [500, 61, 7000, 80, 13]
    tables = [EmbeddingBagConfig(name=f"t{i}", embedding_dim=dims, num_embeddings=h, feature_names=[keys[i]]) for i, h in enumerate(hashes)]
    ebc = EmbeddingBagCollection(tables, device=torch.device("meta"))
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.05})
    model = DLRMTrain(DLRM(ebc, 13, [32, dims], [64, 1], dense_device=device))
    gens = {f"t{i}": sp.table_wise(rank=0, compute_kernel=kernel) for i in range(5)}
    plan = sp.construct_module_sharding_plan(ebc, gens, sharder=EmbeddingBagCollectionSharder(), world_size=1, local_size=1, device_type="cuda")
    dmp = DistributedModelParallel(model, device=device, plan=ShardingPlan({"model.sparse_arch.embedding_bag_collection": plan}), sharders=[EmbeddingBagCollectionSharder()])
    with torch.no_grad():
        g = torch.Generator().manual_seed(11 + seed)
        for _, w, _st, _tbe in dmp.module.model.sparse_arch.embedding_bag_collection._engine.local_shard_views():
            w.copy_((torch.randn(w.shape, generator=g) * 0.1).to(w.device))
        for p in dmp.parameters():
            if p.requires_grad:
                p.copy_((torch.randn(p.shape, generator=g) * 0.05).to(p.device))
    dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(dmp.named_parameters())), lambda p: torch.optim.SGD(p, lr=0.05))
    return dmp, CombinedOptimizer([dmp.fused_optimizer, dense_opt]), keys, hashes


def _weights(dmp):
    eng = dmp.module.model.sparse_arch.embedding_bag_collection._engine
    emb = torch.cat([w.flatten().float() for _, w, _s, _t in eng.local_shard_views()])
    dense = torch.cat([p.detach().flatten().float() for p in dmp.parameters() if p.requires_grad])
    return emb, dense


@pytest.mark.parametrize("pipeline", ["sparse", "sparse_late", "base", "lite", "fused", "prefetch", "semi_sync_first_steps"])
def test_cuda_pipeline_matches_plain_loop(pipeline):
    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.parallel import train_pipeline as tp

    dev = torch.device("cuda:0")
    dmp_a, opt_a, keys, hashes = _build(dev)
    dmp_b, opt_b, _, _ = _build(dev)
    n = 7  # > 3: the plane's id slots wrap around, graph replays kick in
    ds = RandomRecDataset(keys, 64, hash_sizes=hashes, ids_per_features=[1, 3, 2, 4, 1], min_ids_per_features=[1, 0, 0, 0, 1], num_dense=13, manual_seed=5,
                          num_generated_batches=n, num_batches=n, pin_memory=True)
    host = list(iter(ds))
    ref = []
    for b in host:
        b = b.to(dev)
        opt_a.zero_grad()
        loss, _ = dmp_a(b)
        loss.backward()
        opt_a.step()
        ref.append(float(loss))
    kw = {}
    cls = {"sparse": tp.TrainPipelineSparseDist, "sparse_late": tp.TrainPipelineSparseDist, "base": tp.TrainPipelineBase, "lite": tp.TrainPipelineSparseDistLite,
           "fused": tp.TrainPipelineFusedSparseDist, "prefetch": tp.PrefetchTrainPipelineSparseDist, "semi_sync_first_steps": tp.TrainPipelineSemiSync}[pipeline]
    if pipeline == "sparse_late":
        kw = {"data_dist_after_forward": True, "enqueue_batch_after_forward": True}
    pipe = cls(dmp_b, opt_b, dev, **kw)
    it = iter(host)
    got = []
    while True:
        try:
            out = pipe.progress(it)
        except StopIteration:
            break
        got.append(float(out[0]))
    torch.cuda.synchronize()
    assert len(got) == n
    if pipeline == "semi_sync_first_steps":
        # semi-synchronous training looks embeddings up one step early (stale by one update): same first loss, a finite run, weights move
        assert got[0] == pytest.approx(ref[0], rel=1e-4) and all(x == x for x in got)
        return
    assert got == pytest.approx(ref, rel=2e-4, abs=1e-5)
    (ea, da), (eb, db) = _weights(dmp_a), _weights(dmp_b)
    torch.testing.assert_close(eb, ea, rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(db, da, rtol=2e-3, atol=1e-5)


def test_cuda_prefetch_pipeline_with_uvm_cached_tables():
    """fused_uvm_caching tables: the prefetch stage stages the next batch's rows into the HBM cache on its own stream."""
    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.parallel import train_pipeline as tp

    dev = torch.device("cuda:0")
    dmp_a, opt_a, keys, hashes = _build(dev, kernel="fused_uvm_caching")
    dmp_b, opt_b, _, _ = _build(dev, kernel="fused_uvm_caching")
    n = 5
    ds = RandomRecDataset(keys, 32, hash_sizes=hashes, ids_per_features=[1, 3, 2, 4, 1], num_dense=13, manual_seed=8, num_generated_batches=n, num_batches=n, pin_memory=True)
    host = list(iter(ds))
    ref = []
    for b in host:
        b = b.to(dev)
        opt_a.zero_grad()
        loss, _ = dmp_a(b)
        loss.backward()
        opt_a.step()
        ref.append(float(loss))
    pipe = tp.PrefetchTrainPipelineSparseDist(dmp_b, opt_b, dev)
    it = iter(host)
    got = []
    while True:
        try:
            got.append(float(pipe.progress(it)[0]))
        except StopIteration:
            break
    assert got == pytest.approx(ref, rel=2e-4, abs=1e-5)
