"""TrainPipelineSparseDist / Base vs a plain training loop on 2 CPU ranks: identical losses and
weights step by step (methodology of the reference's train_pipeline/tests/test_train_pipelines.py)."""
import copy

import pytest
import torch

from torchrec_b200.utils.multiprocess import run_multi_process


def _build(ctx, sharding, seed=0):
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
    keys = [f"f{i}" for i in range(4)]
    hashes = [50, 60, 70, 80]
    tables = [EmbeddingBagConfig(name=f"t{i}", embedding_dim=8, num_embeddings=h, feature_names=[keys[i]]) for i, h in enumerate(hashes)]
    ebc = EmbeddingBagCollection(tables)
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.05})
    model = DLRMTrain(DLRM(ebc, 5, [16, 8], [16, 1]))
    W = ctx.world_size
    gens = {f"t{i}": (sp.table_wise(rank=i % W) if sharding == "tw" else sp.row_wise()) for i in range(4)}
    plan = sp.construct_module_sharding_plan(ebc, gens, sharder=EmbeddingBagCollectionSharder(), world_size=W, local_size=W, device_type="cpu")
    dmp = DistributedModelParallel(model, device=torch.device("cpu"), plan=ShardingPlan({"model.sparse_arch.embedding_bag_collection": plan}),
                                   sharders=[EmbeddingBagCollectionSharder()])
    dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(dmp.named_parameters())), lambda p: torch.optim.SGD(p, lr=0.1))
    opt = CombinedOptimizer([dmp.fused_optimizer, dense_opt])
    return dmp, opt, keys, hashes


def _run(ctx, sharding: str, pipeline: str):
    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.parallel import train_pipeline as tp

    dmp_a, opt_a, keys, hashes = _build(ctx, sharding)
    dmp_b, opt_b, _, _ = _build(ctx, sharding)
    dmp_b.load_state_dict(dmp_a.state_dict())
    ds = RandomRecDataset(keys, 6, hash_sizes=hashes, ids_per_feature=3, min_ids_per_feature=0, num_dense=5, manual_seed=7 + ctx.rank,
                          num_generated_batches=5, num_batches=5)
    batches = list(iter(ds))
    # plain loop
    ref_losses = []
    for b in batches:
        opt_a.zero_grad()
        loss, _ = dmp_a(b)
        loss.backward()
        opt_a.step()
        ref_losses.append(loss.detach().clone())
    cls = {"sparse": tp.TrainPipelineSparseDist, "base": tp.TrainPipelineBase, "lite": tp.TrainPipelineSparseDistLite,
           "fused": tp.TrainPipelineFusedSparseDist, "prefetch": tp.PrefetchTrainPipelineSparseDist}[pipeline]
    knobs = {"clear_data_dist_inputs": True, "enable_inplace_copy_batch": True} if pipeline in ("sparse", "prefetch") else {}  # the memory knobs change nothing observable
    pipe = cls(dmp_b, opt_b, torch.device("cpu"), **knobs)
    it = iter(batches)
    losses = []
    while True:
        try:
            out = pipe.progress(it)
        except StopIteration:
            break
        losses.append(out[0].clone())
    assert len(losses) == len(ref_losses), (len(losses), len(ref_losses))
    for a, b in zip(losses, ref_losses):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    sa, sb = dmp_a.state_dict(), dmp_b.state_dict()
    for k in sa:
        ta, tb = sa[k], sb[k]
        if hasattr(ta, "local_shards"):
            for x, y in zip(ta.local_shards(), tb.local_shards()):
                torch.testing.assert_close(x.tensor, y.tensor, rtol=1e-5, atol=1e-6)
        else:
            torch.testing.assert_close(ta, tb, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("pipeline", ["sparse", "base", "lite", "fused", "prefetch"])
def test_pipeline_matches_plain_loop_tw(pipeline):
    run_multi_process(_run, world_size=2, backend="gloo", sharding="tw", pipeline=pipeline)


def test_pipeline_matches_plain_loop_rw():
    run_multi_process(_run, world_size=2, backend="gloo", sharding="rw", pipeline="sparse")


def _staged(ctx):
    """StagedTrainPipeline + SparseDataDistUtil (input dist of batch i+1 runs as its own stage) == plain loop."""
    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.parallel import train_pipeline as tp

    dmp_a, opt_a, keys, hashes = _build(ctx, "tw")
    dmp_b, opt_b, _, _ = _build(ctx, "tw")
    dmp_b.load_state_dict(dmp_a.state_dict())
    ds = RandomRecDataset(keys, 6, hash_sizes=hashes, ids_per_feature=3, min_ids_per_feature=0, num_dense=5, manual_seed=3 + ctx.rank, num_generated_batches=5, num_batches=5)
    batches = list(iter(ds))
    ref = []
    for b in batches:
        opt_a.zero_grad()
        loss, _ = dmp_a(b)
        loss.backward()
        opt_a.step()
        ref.append(loss.detach().clone())
    util = tp.SparseDataDistUtil(dmp_b, data_dist_stream=None)
    stages = [tp.PipelineStage("copy", lambda b: b.to(torch.device("cpu")), None),
              tp.PipelineStage("sparse_dist", util.start_sparse_data_dist, None, fill_callback=util.wait_sparse_data_dist)]
    pipe = tp.StagedTrainPipeline(stages)
    it = iter(batches)
    got = []
    while True:
        b = pipe.progress(it)
        if b is None:
            break
        util.wait_sparse_data_dist()
        opt_b.zero_grad()
        loss, _ = dmp_b(b)
        loss.backward()
        opt_b.step()
        got.append(loss.detach().clone())
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        torch.testing.assert_close(a, b)
    util.detach()
    loss, _ = dmp_b(batches[0])      # original forwards restored
    assert torch.isfinite(loss)


def test_staged_pipeline_with_sparse_data_dist_util():
    run_multi_process(_staged, world_size=2, backend="gloo")


def _bounded_queue(ctx):
    """execute_all_batches=False: the queue never grows past two batches, every fetched batch but the in-flight tail trains."""
    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.parallel import train_pipeline as tp

    dmp, opt, keys, hashes = _build(ctx, "tw")
    ds = RandomRecDataset(keys, 6, hash_sizes=hashes, ids_per_feature=3, min_ids_per_feature=0, num_dense=5, manual_seed=3 + ctx.rank,
                          num_generated_batches=12, num_batches=12)
    fetched = [0]

    def counting(it):
        for b in it:
            fetched[0] += 1
            yield b

    pipe = tp.TrainPipelineSparseDist(dmp, opt, torch.device("cpu"), execute_all_batches=False)
    it = counting(iter(ds))
    trained, sizes = 0, []
    while True:
        try:
            pipe.progress(it)
        except StopIteration:
            break
        trained += 1
        sizes.append(len(pipe.batches))
    assert max(sizes) <= 2, sizes
    assert fetched[0] == 12
    assert trained >= 10, (trained, sizes)  # only the batches still in flight when the stream ends are dropped


def test_sparse_dist_queue_is_bounded_without_execute_all_batches():
    run_multi_process(_bounded_queue, world_size=2, backend="gloo")


def _run_eval_fused(ctx):
    """EvalPipelineFusedSparseDist: outputs equal the plain eval forward batch by batch, weights untouched, a detached model is
    re-attached by progress."""
    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.parallel import train_pipeline as tp

    dmp, opt, keys, hashes = _build(ctx, "tw")
    ds = RandomRecDataset(keys, 6, hash_sizes=hashes, ids_per_feature=3, min_ids_per_feature=0, num_dense=5, manual_seed=3 + ctx.rank, num_generated_batches=4, num_batches=4)
    batches = list(iter(ds))
    dmp.eval()
    with torch.no_grad():
        want = [dmp(b) for b in batches]
    def local(v):
        if hasattr(v, "local_shards"):
            return v.local_shards()[0].tensor if v.local_shards() else None
        return v if isinstance(v, torch.Tensor) else None

    before = {k: (local(v).clone() if local(v) is not None else None) for k, v in dmp.state_dict().items()}
    pipe = tp.EvalPipelineFusedSparseDist(dmp, opt, torch.device("cpu"))
    it = iter(batches)
    got = [pipe.progress(it), pipe.progress(it)]
    pipe.detach()
    assert not pipe._model_attached
    while True:
        try:
            got.append(pipe.progress(it))  # re-attaches
        except StopIteration:
            break
    assert pipe._model_attached and len(got) == len(want)
    for g, w in zip(got, want):
        g_pred, w_pred = g[1], w[1][1]  # the model returns (loss, (loss, logits, labels)); the pipeline hands out the inner tuple
        torch.testing.assert_close(g_pred, w_pred, rtol=1e-5, atol=1e-6)
        assert not g_pred.requires_grad
    for k, v in dmp.state_dict().items():
        cur = local(v)
        if cur is not None and before[k] is not None:
            assert torch.equal(cur, before[k]), k


def test_eval_pipeline_fused_sparse_dist():
    run_multi_process(_run_eval_fused, world_size=2, backend="gloo")


def test_future_deque_and_postproc_context_switch():
    from concurrent.futures import ThreadPoolExecutor

    from torchrec_b200.parallel.train_pipeline.pipeline_context import CPUEmbeddingTrainPipelineContext, TrainPipelineContext
    from torchrec_b200.parallel.train_pipeline.train_pipelines import TorchCompileConfig
    from torchrec_b200.parallel.train_pipeline.utils import FutureDeque, get_h2d_func, use_context_for_postprocs

    with ThreadPoolExecutor(1) as ex:
        q = FutureDeque([ex.submit(lambda: 1), 2, ex.submit(lambda: 3), ex.submit(lambda: 4)])
        assert q[0] == 1 and q[0] == 1 and q.pop() == 4 and q.popleft() == 1 and list(q)[0] == 2 and q[1] == 3

    class P:
        def __init__(self):
            self.ctx = "cur"

        def get_context(self):
            return self.ctx

        def set_context(self, c):
            self.ctx = c

    ps = [P(), P()]
    nxt = TrainPipelineContext()
    with use_context_for_postprocs(ps, nxt):
        assert all(p.ctx is nxt for p in ps)
    assert all(p.ctx == "cur" for p in ps)
    c = CPUEmbeddingTrainPipelineContext(dense_gpu_device="cuda:0")
    c.gpu_embedding_outputs["ebc"] = torch.zeros(1)
    assert c.dense_gpu_device == "cuda:0" and TorchCompileConfig().compile_on_iter == 3
    assert get_h2d_func(torch.ones(2), torch.device("cpu")).tolist() == [1.0, 1.0]


def _run_mixed_model_pipeline(ctx):
    """Pooled + sequence collections in one model: the sparse-dist pipeline pipelines both sharded modules and matches the plain loop."""
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig, EmbeddingConfig
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.keyed import CombinedOptimizer, KeyedOptimizerWrapper
    from torchrec_b200.optim.optimizers import in_backward_optimizer_filter
    from torchrec_b200.parallel import train_pipeline as tp
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.test_utils.model_input import ModelInput
    from torchrec_b200.parallel.test_utils.test_model import TestMixedEmbeddingSparseArch

    tables = [EmbeddingBagConfig(name="b0", embedding_dim=8, num_embeddings=30, feature_names=["fb0"]), EmbeddingBagConfig(name="b1", embedding_dim=8, num_embeddings=20, feature_names=["fb1"]),
              EmbeddingConfig(name="s0", embedding_dim=4, num_embeddings=25, feature_names=["fs0"])]

    def build():
        torch.manual_seed(0)
        m = TestMixedEmbeddingSparseArch(tables, num_float_features=5)
        apply_optimizer_in_backward(torch.optim.SGD, list(m.ebc.parameters()) + list(m.ec.parameters()), {"lr": 0.1})
        dmp = DistributedModelParallel(m, device=torch.device("cpu"))
        dense = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(dmp.named_parameters())), lambda p: torch.optim.SGD(p, lr=0.05))
        return dmp, CombinedOptimizer([dmp.fused_optimizer, dense])

    a, opt_a = build()
    b, opt_b = build()
    b.load_state_dict(a.state_dict())
    g = torch.Generator().manual_seed(10 + ctx.rank)
    batches = [ModelInput.generate(batch_size=4, tables=tables, num_float_features=5, pooling_avg=3, generator=g) for _ in range(4)]
    ref = []
    for x in batches:
        opt_a.zero_grad()
        loss, _ = a(x)
        loss.backward()
        opt_a.step()
        ref.append(loss.detach().clone())
    pipe = tp.TrainPipelineSparseDist(b, opt_b, torch.device("cpu"))
    it, got = iter(batches), []
    while True:
        try:
            got.append(pipe.progress(it))
        except StopIteration:
            break
    assert len(pipe._pipelined_modules) == 2, [type(m).__name__ for m in pipe._pipelined_modules]
    # the pipeline returns the model output (loss, prediction): compare predictions of the last step and the weights
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if hasattr(sa[k], "local_shards"):
            for x, y in zip(sa[k].local_shards(), sb[k].local_shards()):
                torch.testing.assert_close(x.tensor, y.tensor, rtol=1e-5, atol=1e-6)
        else:
            torch.testing.assert_close(sa[k], sb[k], rtol=1e-5, atol=1e-6)
    assert len(got) == 4


def test_mixed_pooled_and_sequence_model_pipeline():
    run_multi_process(_run_mixed_model_pipeline, world_size=2, backend="gloo")
