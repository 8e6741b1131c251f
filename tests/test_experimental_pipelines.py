"""Experimental pipelines + HybridEvalDMP (CPU)."""
import torch
from torch import nn

from torchrec_b200.datasets.random import RandomRecDataset


def _model(device=torch.device("cpu"), cls=None, **kw):
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

    torch.manual_seed(0)
    keys = ["a", "b", "c"]
    tables = [EmbeddingBagConfig(name=f"t_{k}", embedding_dim=8, num_embeddings=50, feature_names=[k]) for k in keys]
    ebc = EmbeddingBagCollection(tables, device=torch.device("meta"))
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.05})
    model = DLRMTrain(DLRM(ebc, 13, [16, 8], [16, 1], dense_device=device))
    plan = sp.construct_module_sharding_plan(ebc, {t.name: sp.table_wise(rank=0) for t in tables}, sharder=EmbeddingBagCollectionSharder(), world_size=1, local_size=1, device_type="cpu")
    dmp = (cls or DistributedModelParallel)(model, device=device, plan=ShardingPlan({"model.sparse_arch.embedding_bag_collection": plan}), sharders=[EmbeddingBagCollectionSharder()], **kw)
    dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(dmp.named_parameters())), lambda p: torch.optim.SGD(p, lr=0.05))
    ds = RandomRecDataset(keys, 16, hash_sizes=[50] * 3, ids_per_features=[2] * 3, num_dense=13, manual_seed=1, num_generated_batches=6)
    return dmp, CombinedOptimizer([dmp.fused_optimizer, dense_opt]), ds.batch_generator._generated_batches


def _weights(dmp):
    return torch.cat([t.weights.detach().flatten().clone() for t in dmp.module.model.sparse_arch.embedding_bag_collection.engine._tbes])


def test_stash_and_bwd_injection_pipelines_match_plain():
    from torchrec_b200.parallel.memory_stashing import MemoryStashingManager
    from torchrec_b200.parallel.train_pipeline import TrainPipelineSparseDist
    from torchrec_b200.parallel.train_pipeline.backward_injection import InjectionSite, InjectionTargetType
    from torchrec_b200.parallel.train_pipeline.experimental_pipelines import (TrainPipelinePrefetchEMS, TrainPipelineSparseDistBwdOpt,
                                                                                TrainPipelineSparseDistEmbStash, TrainPipelineSparseDistOptStash, TrainPipelineSparseDistT)

    ref_dmp, ref_opt, batches = _model()
    ref = TrainPipelineSparseDist(ref_dmp, ref_opt, torch.device("cpu"))
    it = iter(batches)
    ref_losses = [float(ref.progress(it)[0]) for _ in range(5)]
    site = InjectionSite(fqn="model.over_arch", target_type=InjectionTargetType.PARAM_GRAD)
    for cls, kw in ((TrainPipelineSparseDistOptStash, {"site": site}), (TrainPipelineSparseDistEmbStash, {}), (TrainPipelineSparseDistT, {}),
                    (TrainPipelinePrefetchEMS, {}), (TrainPipelineSparseDistOptStash, {}),
                    (TrainPipelineSparseDistBwdOpt, {"site": site, "injected_work": lambda p: None})):
        MemoryStashingManager.reset()
        dmp, opt, _ = _model()
        pipe = cls(dmp, opt, torch.device("cpu"), **kw)
        it = iter(batches)
        losses = [float(pipe.progress(it)[0]) for _ in range(5)]
        torch.testing.assert_close(torch.tensor(losses), torch.tensor(ref_losses), rtol=1e-5, atol=1e-6)
        if cls is TrainPipelineSparseDistOptStash:
            assert MemoryStashingManager.stashed_bytes() > 0  # state parked on the host between steps
        if cls is TrainPipelineSparseDistBwdOpt:
            assert pipe.injected_calls == 5
        MemoryStashingManager.reset()
        torch.testing.assert_close(_weights(dmp), _weights(ref_dmp), rtol=1e-5, atol=1e-6)


def test_train_eval_hybrid_pipeline():
    from torchrec_b200.parallel.train_pipeline.experimental_pipelines import TrainEvalHybridPipelineBase

    dmp, opt, batches = _model()
    flags = [False, True, False, True, False]
    for b, f in zip(batches, flags):
        b.labels[0] = 7 if f else b.labels[0]  # the marker must survive Batch.to(device): carried in the data itself
    pipe = TrainEvalHybridPipelineBase(dmp, opt, torch.device("cpu"), is_eval_batch=lambda b: int(b.labels[0]) == 7)
    it = iter(batches[:5])
    w_hist = []
    for _ in range(5):
        pipe.progress(it)
        w_hist.append(_weights(dmp))
    assert not torch.equal(w_hist[0], w_hist[2]) and torch.equal(w_hist[0], w_hist[1]) and torch.equal(w_hist[2], w_hist[3])
    assert dmp.training


def test_hybrid_eval_dmp_and_cpu_sparse_eval():
    from torchrec_b200.parallel.model_parallel import HybridEvalDMP
    from torchrec_b200.parallel.train_pipeline.experimental_pipelines import EvalPipelineCPUSparse

    dmp, _, batches = _model(cls=HybridEvalDMP)
    assert not dmp.training
    dmp.to(torch.device("cpu"))
    m = dmp.module.model

    def sparse(batch):
        return m.sparse_arch(batch.sparse_features)

    def dense(batch, pooled):
        d = m.dense_arch(batch.dense_features)
        return m.over_arch(m.inter_arch(dense_features=d, sparse_features=pooled))

    pipe = EvalPipelineCPUSparse(sparse, dense, torch.device("cpu"))
    it = iter(batches[:3])
    outs = [pipe.progress(it) for _ in range(3)]
    with torch.no_grad():
        ref = [m(b.dense_features, b.sparse_features) for b in batches[:3]]
    for o, r in zip(outs, ref):
        torch.testing.assert_close(o, r)


def test_data_loading_thread_postproc_and_pt2_pipeline():
    import torch
    from torch import nn

    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.models.dlrm import DLRM, DLRMTrain
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.parallel.train_pipeline import DataLoadingThread, PipelinedPostproc, TrainPipelineContext, TrainPipelinePT2

    dev = torch.device("cpu")
    keys = ["f0", "f1"]
    ds = RandomRecDataset(keys=keys, batch_size=8, hash_size=50, ids_per_feature=2, num_dense=4, manual_seed=0, num_batches=5)
    t = DataLoadingThread(dev, iter(ds), queue_size=2)
    t.start()
    got = []
    while True:
        b = t.get(timeout=10)
        if b is None:
            break
        got.append(b)
    assert len(got) == 5 and got[0].dense_features.shape == (8, 4) and t.get(timeout=1) is None
    t.join(timeout=5)

    calls = []

    class Remap(nn.Module):
        def forward(self, x):
            calls.append(1)
            return x * 2

    ctx = TrainPipelineContext(version=1)
    pp = PipelinedPostproc(Remap(), "model.remap", ctx)
    x = torch.ones(3)
    assert torch.equal(pp(x), x * 2) and torch.equal(pp(x), x * 2) and len(calls) == 1      # second call: cached for this batch context
    pp.set_context(TrainPipelineContext(version=1))
    pp(x)
    assert len(calls) == 2

    tables = [EmbeddingBagConfig(name=f"t{i}", embedding_dim=8, num_embeddings=50, feature_names=[k]) for i, k in enumerate(keys)]
    model = DLRMTrain(DLRM(EmbeddingBagCollection(tables), dense_in_features=4, dense_arch_layer_sizes=[8, 8], over_arch_layer_sizes=[8, 1]))
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    seen = {"pre": 0, "compiled": 0, "transformed": 0}

    def compile_fn(m):
        seen["compiled"] += 1
        return m

    def transform(b):
        seen["transformed"] += 1
        return b

    pipe = TrainPipelinePT2(model, opt, dev, compile_fn=compile_fn, pre_compile_fn=lambda m: seen.__setitem__("pre", seen["pre"] + 1), input_transformer=transform,
                            num_pre_compile_steps=2)
    it = iter(RandomRecDataset(keys=keys, batch_size=8, hash_size=50, ids_per_feature=2, num_dense=4, manual_seed=1, num_batches=6))
    losses = []
    try:
        while True:
            losses.append(float(pipe.progress(it)[0]))
    except StopIteration:
        pass
    assert len(losses) == 6 and seen["compiled"] == 1 and seen["pre"] == 1 and seen["transformed"] == 6


def test_runtime_forwards_and_arg_info():
    import torch

    from torchrec_b200.parallel.train_pipeline.pipeline_context import EmbeddingTrainPipelineContext, PrefetchTrainPipelineContext
    from torchrec_b200.parallel.train_pipeline.runtime_forwards import (CPUEmbeddingPipelinedForward, EmbeddingPipelinedForward, InSyncEmbeddingPipelinedForward,
                                                                       PrefetchEmbeddingPipelinedForward, PrefetchPipelinedForward)
    from torchrec_b200.parallel.train_pipeline.types import ArgInfo, CallArgs, PipelinePhase, PipelineState
    from torchrec_b200.parallel.types import NoWait

    class FakeSharded:
        def __init__(self):
            self.calls = []

        def compute_and_output_dist(self, ctx, data):
            self.calls.append((ctx, data))
            return data * 10

    m = FakeSharded()
    ctx = EmbeddingTrainPipelineContext(version=1, index=3)
    ctx.embedding_a2a_requests["ebc"] = NoWait(torch.ones(2))
    ctx.module_contexts["ebc"] = "mctx"
    fwd = EmbeddingPipelinedForward("ebc", None, m, ctx)
    assert torch.equal(fwd("ignored"), torch.ones(2)) and not ctx.embedding_a2a_requests and m.calls == []

    ctx2 = EmbeddingTrainPipelineContext(version=1, index=3)
    ctx2.embedding_a2a_requests["ebc"] = NoWait(torch.ones(2))          # computed before the update at step 5 -> stale
    ctx2.input_dist_tensors_requests["ebc"] = NoWait(torch.full((2,), 2.0))
    ctx2.module_contexts["ebc"] = "mctx"
    insync = InSyncEmbeddingPipelinedForward("ebc", None, m, ctx2, fresh_since=lambda: 5)
    assert torch.equal(insync(), torch.full((2,), 20.0)) and len(m.calls) == 1
    ctx2.index = 7
    ctx2.embedding_a2a_requests["ebc"] = NoWait(torch.ones(2))
    assert torch.equal(insync(), torch.ones(2))                           # fresh enough: precomputed result is used

    ctx3 = EmbeddingTrainPipelineContext(version=1)
    ctx3.embedding_a2a_requests["ebc"] = NoWait({"f": torch.ones(1)})
    cpu = CPUEmbeddingPipelinedForward("ebc", None, m, ctx3, device=torch.device("cpu"))
    assert cpu()["f"].device.type == "cpu"

    pctx = PrefetchTrainPipelineContext(version=1)
    pctx.module_input_post_prefetch["ebc"] = torch.full((2,), 3.0)
    pctx.module_contexts_post_prefetch["ebc"] = "pm"
    assert torch.equal(PrefetchPipelinedForward("ebc", None, m, pctx)(), torch.full((2,), 30.0)) and m.calls[-1][0] == "pm"
    pctx.module_input_post_prefetch["ebc"] = torch.full((2,), 3.0)
    pctx.module_contexts_post_prefetch["ebc"] = "pm"
    assert torch.equal(PrefetchEmbeddingPipelinedForward("ebc", None, m, pctx)(), torch.full((2,), 30.0))

    class B:
        sparse_features = {"k": [10, 20, 30]}

    info = ArgInfo.from_path([("attr", "sparse_features"), ("item", "k"), ("item", 1)])
    assert info.process(B()) == 20
    args, kwargs = CallArgs([info], {"x": ArgInfo.from_path([("attr", "sparse_features")])}).build_args_kwargs(B())
    assert args == [20] and kwargs["x"] is B.sparse_features
    assert str(PipelineState.CALL_FWD) == "CALL_FWD" and PipelinePhase.FORWARD.value == "forward"


def test_fx_tracing_discovers_pipelineable_modules_and_recipes():
    """``train_pipeline/tracing.py``: shallow fx trace with sharded modules as leaves; inputs walked back to the batch as ArgInfo steps."""
    import dataclasses

    import torch
    from torch import nn

    from torchrec_b200.parallel.train_pipeline.pipeline_stage import PipelineStage, SparseDataDistUtil  # noqa: F401
    from torchrec_b200.parallel.train_pipeline.postproc import NoOpStream, PipelinedPostproc  # noqa: F401
    from torchrec_b200.parallel.train_pipeline.tracing import ArgInfoStepFactory, Tracer, _get_leaf_module_names, rewrite_model
    from torchrec_b200.parallel.types import ShardedModule

    class FakeSharded(ShardedModule):
        def __init__(self):
            nn.Module.__init__(self)

        def create_context(self):
            return None

        def input_dist(self, ctx, *a, **k):
            pass

        def compute(self, ctx, x):
            pass

        def output_dist(self, ctx, x):
            pass

        def forward(self, x):
            return x * 2

    class Clamp(nn.Module):
        def forward(self, x):
            return x.clamp(max=3)

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c, self.pp, self.lin = FakeSharded(), FakeSharded(), FakeSharded(), Clamp(), nn.Linear(2, 2)

        def forward(self, batch):
            return self.a(batch.sparse["u"]) + self.b(self.pp(batch.sparse["v"])) + self.c(self.lin(batch.dense)).sum()

    @dataclasses.dataclass
    class B:
        sparse: dict
        dense: torch.Tensor

    m = M()
    assert sorted(_get_leaf_module_names(m)) == ["a", "b", "c", "lin", "pp"]
    ok, bad, postprocs = rewrite_model(m, pipeline_postproc=True)
    assert [i.fqn for i in ok] == ["a", "b"] and bad == ["c"] and [type(p).__name__ for p in postprocs] == ["Clamp"]
    b = B({"u": torch.tensor([1.0, 5.0]), "v": torch.tensor([2.0, 9.0])}, torch.ones(2))
    assert ok[0].call_args.build_args_kwargs(b)[0][0].tolist() == [1.0, 5.0]
    assert ok[1].call_args.build_args_kwargs(b)[0][0].tolist() == [2.0, 3.0]          # the postproc runs inside the recipe
    ok2, bad2, _ = rewrite_model(m, pipeline_postproc=False)
    assert [i.fqn for i in ok2] == ["a"] and bad2 == ["b", "c"]
    assert ArgInfoStepFactory.from_scalar(7).process(None) == 7 and ArgInfoStepFactory.get_item("k").process({"k": 1}) == 1
    with NoOpStream() as s:
        s.wait_stream(None)
