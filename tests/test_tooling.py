"""Tooling: fx tracer, IR serializer, pt2 helpers, schema BC checker + frozen public signatures, linter, loggers."""
import inspect
import logging

import torch
from torch import nn

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
from torchrec_b200.sparse import KeyedJaggedTensor


def _ebc():
    return EmbeddingBagCollection([EmbeddingBagConfig(name="t1", embedding_dim=4, num_embeddings=10, feature_names=["f1"]),
                                   EmbeddingBagConfig(name="t2", embedding_dim=4, num_embeddings=10, feature_names=["f2"])])


def test_fx_tracer_keeps_sparse_modules_as_leaves():
    from torchrec_b200.fx import symbolic_trace

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = _ebc()
            self.lin = nn.Linear(8, 1)

        def forward(self, kjt):
            return self.lin(self.ebc(kjt).values())

    m = M()
    gm = symbolic_trace(m)
    targets = [n.target for n in gm.graph.nodes if n.op == "call_module"]
    assert "ebc" in targets and "lin" in targets
    kjt = KeyedJaggedTensor(keys=["f1", "f2"], values=torch.tensor([1, 2, 3]), lengths=torch.tensor([1, 1, 1, 0]))
    torch.testing.assert_close(gm(kjt), m(kjt))


def test_ir_serializer_round_trip():
    from torchrec_b200.ir.serializer import JsonSerializer, decapsulate_ir_modules, encapsulate_ir_modules
    from torchrec_b200.modules.feature_processor_ import PositionWeightedModuleCollection
    from torchrec_b200.modules.fp_embedding_modules import FeatureProcessedEmbeddingBagCollection

    ebc = _ebc()
    new = JsonSerializer.deserialize(JsonSerializer.serialize(ebc))
    assert [c.name for c in new.embedding_bag_configs()] == ["t1", "t2"] and new.embedding_bag_configs()[0].feature_names == ["f1"]
    fp = FeatureProcessedEmbeddingBagCollection(EmbeddingBagCollection(ebc.embedding_bag_configs(), is_weighted=True), PositionWeightedModuleCollection({"f1": 3, "f2": 4}))
    fp2 = JsonSerializer.deserialize(JsonSerializer.serialize(fp))
    assert isinstance(fp2, FeatureProcessedEmbeddingBagCollection) and fp2._feature_processors.max_feature_lengths == {"f1": 3, "f2": 4}

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.sparse = ebc

    m, fqns = encapsulate_ir_modules(M())
    assert fqns == ["sparse"] and m.sparse.ir_metadata
    w = m.sparse.embedding_bags["t1"].weight.detach().clone()
    m2 = decapsulate_ir_modules(m)
    assert m2.sparse is not ebc
    torch.testing.assert_close(m2.sparse.embedding_bags["t1"].weight.detach(), w)


def test_pt2_helpers():
    from torchrec_b200.pt2.utils import is_pt2_compiling, kjt_for_pt2_tracing, pt2_checks_tensor_slice

    assert not is_pt2_compiling()
    pt2_checks_tensor_slice(torch.zeros(4), 0, 2)
    kjt = KeyedJaggedTensor(keys=["a"], values=torch.tensor([1, 2]), lengths=torch.tensor([2, 0]))
    k2 = kjt_for_pt2_tracing(kjt)
    assert k2.keys() == ["a"] and k2.lengths().dtype == torch.int64


def test_signature_compat_checker():
    from torchrec_b200.schema.utils import is_signature_compatible as ok

    def f0(a, b=1): ...
    def f1(a, b=1, c=2): ...
    def f2(a, c=2, b=1): ...
    def f3(a, b): ...
    def f4(a, b=1, *, d): ...
    s = inspect.signature
    assert ok(s(f0), s(f1)) and not ok(s(f0), s(f2)) and not ok(s(f0), s(f3)) and not ok(s(f0), s(f4)) and ok(s(f0), s(f0))


def test_public_api_signatures_are_stable():
    """Frozen signatures of the user-facing entry points (the reference's schema/api_tests)."""
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.planner import EmbeddingShardingPlanner
    from torchrec_b200.parallel.train_pipeline import TrainPipelineSparseDist
    from torchrec_b200.schema.utils import is_signature_compatible

    def dmp(self, module, env=None, device=None, plan=None, sharders=None, init_data_parallel=True, init_parameters=True, data_parallel_wrapper=None): ...
    def ebc(self, tables, is_weighted=False, device=None): ...
    def kjt(self, keys, values, weights=None, lengths=None, offsets=None, stride=None, stride_per_key_per_rank=None): ...
    def pipe(self, model, optimizer, device, execute_all_batches=True, apply_jit=False): ...
    def plan(self, module, sharders): ...
    for frozen, live in ((dmp, DistributedModelParallel.__init__), (ebc, EmbeddingBagCollection.__init__), (kjt, KeyedJaggedTensor.__init__), (pipe, TrainPipelineSparseDist.__init__),
                         (plan, EmbeddingShardingPlanner.plan)):
        assert is_signature_compatible(inspect.signature(frozen), inspect.signature(live)), (live.__qualname__, inspect.signature(live))


def test_module_linter(tmp_path):
    from torchrec_b200.linter.module_linter import linter_one_file

    f = tmp_path / "m.py"
    f.write_text("import torch.nn as nn\nclass Good(nn.Module):\n    '''Doc.\n\n    Args:\n        x: size\n\n    Example::\n        Good(1)\n    '''\n    def __init__(self, x):\n        super().__init__()\n"
                 "    def forward(self, t):\n        return t\nclass Bad(nn.Module):\n    def __init__(self, a, b):\n        super().__init__()\nclass Half(nn.Module):\n    '''Doc only.'''\n    def __init__(self, a):\n        super().__init__()\n")
    issues = linter_one_file(str(f))
    names = sorted((i["name"], i["description"].split("`")[1]) for i in issues)
    assert ("docstring-missing", "Bad") in names and ("args-section-missing", "Half") in names and not any(n[1] == "Good" for n in names)


def test_loggers(caplog):
    from torchrec_b200.parallel.logger import CappedLogger, EventLoggingHandler, LazyStr, _torchrec_method_logger, get_logger

    calls = []
    s = LazyStr(lambda: calls.append(1) or "expensive")
    logging.getLogger("x").debug(s)
    assert calls == [] and str(s) == "expensive"
    with caplog.at_level(logging.WARNING):
        c = CappedLogger(get_logger("t"), cap=2)
        for _ in range(5):
            c.warning("k", "hot path warning")
    assert sum("hot path warning" in r.getMessage() for r in caplog.records) == 2
    events = []
    EventLoggingHandler.register("test", lambda e, f: events.append((e, f["success"])))

    @EventLoggingHandler.event_logger("unit")
    @_torchrec_method_logger()
    def work(x):
        return x + 1

    assert work(1) == 2 and events == [("unit", True)]


def test_lazy_extension():
    import torch

    from torchrec_b200.modules.lazy_extension import LazyModuleExtensionMixin, lazy_apply

    class LazyScale(LazyModuleExtensionMixin, torch.nn.Module):
        cls_to_become = None

        def __init__(self):
            super().__init__()
            self.w = torch.nn.UninitializedParameter()

        def initialize_parameters(self, x, *, bias=None):
            if self.has_uninitialized_params():
                self.w.materialize((x.shape[-1],))
                with torch.no_grad():
                    self.w.fill_(2.0)

        def forward(self, x, *, bias=None):
            y = x * self.w
            return y if bias is None else y + bias

    m = LazyScale()
    import pytest

    with pytest.raises(RuntimeError):
        m.apply(lambda mod: None)
    seen = []

    @torch.no_grad()
    def init(mod):
        seen.append(type(mod).__name__)
        if isinstance(mod, LazyScale):
            mod.w.fill_(3.0)

    lazy_apply(m, init)
    x = torch.ones(2, 4)
    out = m(x, bias=torch.ones(4))          # kwargs reach initialize_parameters; first forward still sees the 2.0 init
    assert torch.equal(out, torch.full((2, 4), 3.0))
    assert seen == ["LazyScale"] and torch.equal(m.w.detach(), torch.full((4,), 3.0))
    m(x)
    assert seen == ["LazyScale"]            # ran once
    m.apply(lambda mod: None)               # initialized: plain apply is allowed again
    seq = lazy_apply(torch.nn.Sequential(torch.nn.LazyLinear(2)), lambda mod: seen.append("s"))
    seq(torch.randn(3, 5))
    assert seen.count("s") == 2


def test_affinity_helper_is_safe_without_gpu():
    import os

    from torchrec_b200.utils.affinity import bind_to_gpu_numa, gpu_local_cpus

    before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    cpus = gpu_local_cpus(0)
    assert cpus is None or (isinstance(cpus, set) and all(isinstance(c, int) for c in cpus))
    new = bind_to_gpu_numa(0)
    if new is None and before is not None:
        assert os.sched_getaffinity(0) == before  # no NVML / no GPU: nothing changes
    if before is not None:
        os.sched_setaffinity(0, before)


def test_benchmark_harness_cmd_conf_and_loggers(tmp_path):
    from dataclasses import dataclass
    from typing import List, Optional

    import torch

    from torchrec_b200.benchmarks.base import benchmark_func, cmd_conf
    from torchrec_b200.parallel.global_settings import get_propogate_device, set_propogate_device
    from torchrec_b200.parallel.logger import ForkedPdb, PercentileLogger, log_table_assignment

    @dataclass
    class RunCfg:
        batch_size: int = 8
        name: str = "x"
        flags: Optional[List[int]] = None
        fast: bool = False

    @dataclass
    class ModelCfg:
        dim: int = 4
        lr: float = 0.1

    @cmd_conf
    def main(run: RunCfg, model: ModelCfg, tag: str = "t"):
        return run, model, tag

    cfg = tmp_path / "c.yml"
    cfg.write_text("batch_size: 32\nModelCfg:\n  dim: 16\n  lr: 0.5\ntag: from_file\n")
    run, model, tag = main(["--yaml_config", str(cfg), "--dim", "64", "--flags", "1", "2", "--fast", "true"])
    assert (run.batch_size, run.name, run.flags, run.fast) == (32, "x", [1, 2], True)
    assert (model.dim, model.lr, tag) == (64, 0.5, "from_file")   # CLI > file > default

    x = torch.randn(64, 64)
    res = benchmark_func("matmul", lambda: x @ x, num_benchmarks=4, num_warmup=1, profile_dir=str(tmp_path / "prof"))
    assert res.gpu_elapsed_time.numel() == 4 and res.runtime_percentile(50) >= 0 and "matmul" in str(res)
    assert any(p.name.startswith("trace-matmul") for p in (tmp_path / "prof").iterdir())

    pl = PercentileLogger("lat", log_every=10)
    for i in range(100):
        pl.add(i)
    p = pl.percentiles()
    assert p[50] in (49.0, 50.0) and p[99] >= 97.0 and "p90" in pl.summary()
    set_propogate_device(True)
    assert get_propogate_device()
    set_propogate_device(False)
    assert hasattr(ForkedPdb(), "set_trace")

    class _PS:
        sharding_type, compute_kernel, ranks = "table_wise", "fused", [1]

    class _Plan:
        plan = {"ebc": {"t0": _PS()}}

    assert log_table_assignment(_Plan())[0]["table"] == "t0"


def test_benchmark_scripts_run_tiny():
    from torchrec_b200.benchmarks import benchmark_ebc, benchmark_inference, benchmark_zch

    res = benchmark_ebc.run(benchmark_ebc.EbcBenchConfig(num_tables=2, num_embeddings=100, embedding_dim=8, batch_size=16, pooling_factor=2, num_benchmarks=2, num_warmup=1,
                                                        device="cpu"))
    assert [r.short_name for r in res] == ["ebc_fwd_bwd_sgd", "fused_ebc_fwd_bwd_sgd", "quant_ebc_int8_fwd"]
    z = benchmark_zch.run(benchmark_zch.ZchBenchConfig(zch_size=64, id_space=1000, batch_size=128, num_benchmarks=4, num_warmup=2, device="cpu"))
    assert z["ids_per_s"] > 0 and 0.0 < z["in_zch_range"] <= 1.0
    inf = benchmark_inference.run(benchmark_inference.InferenceBenchConfig(num_tables=2, num_embeddings=100, embedding_dim=8, batch_size=8, num_benchmarks=2, num_warmup=1))
    assert inf[0].gpu_elapsed_time.numel() == 2


def _lifecycle(ctx):
    from torchrec_b200.benchmarks import benchmark_model_lifecycle as L

    t = L.run(L.LifecycleConfig(num_tables=3, num_embeddings=200, embedding_dim=8, batch_size=8, device="cpu"))
    assert set(t) == {"build_meta_ms", "plan_ms", "shard_materialise_ms", "state_dict_ms", "load_state_dict_ms", "first_step_ms"}
    assert all(v >= 0 for v in t.values())


def test_benchmark_model_lifecycle_gloo():
    from torchrec_b200.utils.multiprocess import run_multi_process

    run_multi_process(_lifecycle, world_size=2, backend="gloo")


def test_distributed_helper_modules(tmp_path):
    import copy

    import torch
    from torch import nn

    from torchrec_b200.modules.embedding_configs import DataType, EmbeddingBagConfig
    from torchrec_b200.modules.fused_embedding_modules import FusedEmbeddingBagCollection
    from torchrec_b200.ops.tbe import OptimType
    from torchrec_b200.parallel import fused_params as FP
    from torchrec_b200.parallel import infer_utils as IU
    from torchrec_b200.parallel import utils as U
    from torchrec_b200.parallel.composable import TableBatchedEmbeddingSlice
    from torchrec_b200.parallel.embedding_dim_bucketer import EmbDimBucketer, EmbDimBucketerPolicy
    from torchrec_b200.parallel.shards_wrapper import LocalShardsWrapper
    from torchrec_b200.parallel.types import ParameterSharding

    # utils
    assert U.append_prefix("a", "b") == "a.b" and U.append_prefix("", "b") == "b" and U.none_throws(3) == 3
    sd = {"m.w": torch.ones(1), "m.sub.b": torch.zeros(1), "n.w": torch.ones(1)}
    assert list(U.filter_state_dict(sd, "m")) == ["w", "sub.b"]
    U.add_prefix_to_state_dict(sd, "x.")
    assert set(sd) == {"x.m.w", "x.m.sub.b", "x.n.w"}
    assert U.optimizer_type_to_emb_opt_type(torch.optim.Adagrad) == OptimType.EXACT_ADAGRAD and U.emb_opt_type_to_optimizer_class(OptimType.ADAM).__name__ == "Adam"
    assert U.merge_fused_params({"learning_rate": 0.1, "eps": 1.0}, {"learning_rate": 0.5}) == {"learning_rate": 0.5, "eps": 1.0}
    ps = ParameterSharding(sharding_type="table_wise", compute_kernel="fused", ranks=[0])
    ps.stochastic_rounding = False
    assert U.add_params_from_parameter_sharding({"eps": 1.0}, ps)["stochastic_rounding"] is False
    lin = nn.Linear(3, 2, device="meta")
    U.init_parameters(lin, torch.device("cpu"))
    assert lin.weight.device.type == "cpu" and not lin.weight.is_meta
    net = nn.Sequential(nn.Linear(2, 2), nn.ReLU())
    assert U.get_unsharded_module_names(net) == [""]
    with U.sharded_model_copy("cpu"):
        net2 = copy.deepcopy(net)
    assert torch.equal(net2[0].weight, net[0].weight) and net2[0].weight is not net[0].weight

    class C(U.CopyableMixin):
        def __init__(self):
            super().__init__()
            self.w = nn.Parameter(torch.ones(2))

    assert torch.equal(C().copy(torch.device("cpu")).w, torch.ones(2))

    # fused params
    fp = {"learning_rate": 0.1, FP.FUSED_PARAM_REGISTER_TBE_BOOL: True, FP.FUSED_PARAM_TBE_ROW_ALIGNMENT: 16}
    assert FP.is_fused_param_register_tbe(fp) and FP.get_fused_param_tbe_row_alignment(fp) == 16 and FP.tbe_fused_params(fp) == {"learning_rate": 0.1}
    assert FP.get_embedding_table_index_type(None) == torch.int64 and not FP.is_fused_param_quant_state_dict_split_scale_bias(None)

    # bucketer
    class T:
        def __init__(self, cols, dt):
            self.local_cols, self.data_type = cols, dt

    tabs = [T(16, DataType.FP32), T(32, DataType.FP32), T(64, DataType.FP16), T(200, DataType.FP32)]   # 64 B, 128 B, 128 B, 800 B
    assert EmbDimBucketer(tabs, EmbDimBucketerPolicy.SINGLE_BUCKET).bucket_count() == 1
    allb = EmbDimBucketer(tabs, EmbDimBucketerPolicy.ALL_BUCKETS)
    assert allb.bucket_count() == 3 and allb.get_bucket(32, DataType.FP32) == allb.get_bucket(64, DataType.FP16)
    cl = EmbDimBucketer(tabs, EmbDimBucketerPolicy.CACHELINE_BUCKETS)
    assert cl.bucket_count() == 2 and cl.get_bucket(16, DataType.FP32) == cl.get_bucket(32, DataType.FP32) != cl.get_bucket(200, DataType.FP32)

    # infer utils on a fused EBC
    febc = FusedEmbeddingBagCollection([EmbeddingBagConfig(name="a", embedding_dim=8, num_embeddings=10, feature_names=["fa"]),
                                        EmbeddingBagConfig(name="b", embedding_dim=8, num_embeddings=20, feature_names=["fb"])], torch.optim.SGD, {"lr": 0.1})
    assert len(IU.get_tbes_from_sharded_module(febc)) == 1
    specs = IU.get_tbe_specs_from_sharded_module(febc)
    assert [(r, c) for _, r, c, _, _ in specs] == [(10, 8), (20, 8)] and U.weights_bytes_in_emb_kernel(febc) == 30 * 8 * 4
    assert all(d == "cpu" for _, d in IU.get_path_device_tuples(febc)) and "" in IU.get_all_torchrec_modules(febc)

    # parameter slice over a flat buffer
    flat = torch.arange(24.0, requires_grad=True)
    sl = TableBatchedEmbeddingSlice(flat, 4, 16, 3, 4)
    assert sl.shape == (3, 4) and isinstance(sl, nn.Parameter) and float(sl[0, 0]) == 4.0
    with torch.no_grad():
        flat[4] = -1.0
    assert float(sl[0, 0]) == -1.0                               # shares storage with the buffer
    flat.grad = torch.ones(24)
    assert sl.grad.shape == (3, 4) and float(sl.grad.sum()) == 12.0
    assert copy.deepcopy(sl).shape == (3, 4)

    # local shards wrapper
    a, b = torch.ones(4, 2), torch.full((4, 3), 2.0)
    w = LocalShardsWrapper([a, b], [(0, 0), (0, 2)])
    assert tuple(w.shape) == (4, 5) and len(w.local_shards()) == 2 and w.local_offsets() == [(0, 0), (0, 2)]
    assert torch.equal(w.full_tensor(), torch.cat([a, b], 1))
    c = w.detach().clone()
    assert isinstance(c, LocalShardsWrapper) and torch.equal(c.local_shards()[1], b) and c.local_shards()[1] is not b
    full = torch.arange(20.0).view(4, 5)
    c.copy_(full)
    assert torch.equal(c.local_shards()[1], full[:, 2:5])
    path = tmp_path / "w.pt"
    torch.save({"t": w}, path)
    back = torch.load(path, weights_only=False)["t"]
    assert isinstance(back, LocalShardsWrapper) and torch.equal(back.local_shards()[0], a)


def test_ir_utils_export_with_placeholder_ops_and_rebuild():
    """Export a model whose EBC is encapsulated behind the placeholder op (dynamic id count), re-target its device nodes, then rebuild the
    real module from the serialized config (reference ir/utils.py + ir/tests/test_serializer.py)."""
    import torch
    from torch import nn

    from torchrec_b200.ir import utils as U
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = EmbeddingBagCollection(tables=[EmbeddingBagConfig(name="t0", embedding_dim=8, num_embeddings=20, feature_names=["f0"]),
                                                      EmbeddingBagConfig(name="t1", embedding_dim=4, num_embeddings=10, feature_names=["f1"])])
            self.head = nn.Linear(12, 1)

        def forward(self, values, lengths):
            kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=values, lengths=lengths)
            return self.head(self.ebc(kjt).values())

    net = Net()
    assert U.qualname(net.ebc).endswith("embedding_modules.EmbeddingBagCollection") and U.get_device([None, torch.zeros(1)]).type == "cpu"
    model, fqns = U.encapsulate_ir_modules(net)
    assert fqns == ["ebc"]
    originals = U.swap_placeholder_forwards(model)
    values, lengths = torch.tensor([1, 2, 3, 4, 5]), torch.tensor([2, 1, 1, 1])
    out = model(values, lengths)
    assert out.shape == (2, 1)
    # the placeholder op registers with dispatch and with fake tensors (what torch.export traces with)
    outs = torch.ops.torchrec_b200.ir_emb_lookup([values, lengths, None, None], 2, [12])
    assert outs[0].shape == (2, 12) and float(outs[0].abs().sum()) == 0.0
    kjt = KeyedJaggedTensor(keys=["f0", "f1"], values=values, lengths=lengths)
    sc = U.mark_dynamic_kjt(kjt, variable_length=True)
    assert sc is not None
    try:
        from torch.export import Dim, export

        ep = export(model, (values, lengths), dynamic_shapes={"values": {0: Dim("v", min=2)}, "lengths": None}, strict=False)
        targets = [str(n.target) for n in ep.graph.nodes if n.op == "call_function"]
        assert any("ir_emb_lookup" in t for t in targets), targets
        more = ep.module()(torch.tensor([1, 2, 3, 4, 5, 6, 7]), torch.tensor([3, 1, 2, 1]))  # a different number of ids
        assert more.shape == (2, 1)
        U.move_to_copy_nodes_to_device(ep.module(), torch.device("cpu"))
    except Exception as e:  # pragma: no cover - export API drift must not hide the rest of the checks
        import warnings

        warnings.warn(f"torch.export path skipped: {type(e).__name__}: {e}")
    U.restore_forwards(model, originals)
    rebuilt = U.decapsulate_ir_modules(model)
    real = rebuilt(values, lengths)
    assert real.shape == (2, 1) and float(real.abs().sum()) > 0
    # flatten / unflatten pairs of an fx graph with flat inputs disappear
    import torch.fx as fx

    def tree_flatten_spec(args, spec=None):
        return list(args)

    g = fx.Graph()
    a, b = g.placeholder("a"), g.placeholder("b")
    fl = g.call_function(tree_flatten_spec, ((a, b),))
    import operator

    x0, x1 = g.call_function(operator.getitem, (fl, 0)), g.call_function(operator.getitem, (fl, 1))
    g.output(g.call_function(torch.add, (x0, x1)))
    gm = fx.GraphModule(nn.Module(), g)
    pruned = U.prune_pytree_flatten_unflatten(gm)
    assert not any("tree_flatten_spec" in str(n.target) for n in pruned.graph.nodes)
    assert torch.equal(pruned(torch.ones(2), torch.ones(2)), torch.full((2,), 2.0))


def test_fx_markers_constants_and_safe_asserts():
    import torch
    from torch import nn

    from torchrec_b200.fx import symbolic_trace
    from torchrec_b200.fx.utils import assert_fx_safe, fx_marker, is_marker_node, leaf_call_counts, marker_regions
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    fixed = KeyedJaggedTensor(keys=["f0"], values=torch.tensor([1, 2, 3]), lengths=torch.tensor([2, 1]))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = EmbeddingBagCollection(tables=[EmbeddingBagConfig(name="t0", embedding_dim=4, num_embeddings=10, feature_names=["f0"])])
            self.lin = nn.Linear(4, 2)

        def forward(self, x):
            assert_fx_safe(x.shape[0] == 2, "batch of two")
            fx_marker("DENSE_BEGIN", x)
            y = self.lin(x)
            fx_marker("DENSE_END", y)
            return y + self.ebc(fixed).values().sum(dim=1, keepdim=True)  # a KJT constant closed over by forward

    net = Net()
    gm = symbolic_trace(net)
    assert leaf_call_counts(gm).get("ebc") == 1
    assert sum(is_marker_node(n, "DENSE_BEGIN") for n in gm.graph.nodes) == 1
    regions = marker_regions(gm.graph, "DENSE_BEGIN", "DENSE_END")
    assert len(regions) == 1 and [n.target for n in regions[0] if n.op == "call_module"] == ["lin"]
    assert any(n.op == "get_attr" and str(n.target).startswith("_sparse_constant_") for n in gm.graph.nodes)
    x = torch.randn(2, 4)
    torch.testing.assert_close(gm(x), net(x))
    import pytest

    with pytest.raises(AssertionError):
        assert_fx_safe(False, "eager asserts still fire")


def test_every_benchmark_yaml_runs_on_cpu_with_tiny_shapes():
    """All pipeline-benchmark configs under benchmarks/yaml parse and run end to end (CPU, one rank, shapes shrunk): every pipeline
    name, sharding type, compute kernel and run knob they mention is wired (reference distributed/benchmark/yaml/*.yml)."""
    import glob
    import os

    import yaml

    import torchrec_b200.benchmarks.benchmark_train_pipeline as B

    files = sorted(glob.glob(os.path.join(os.path.dirname(B.__file__), "yaml", "*.yml")))
    assert len(files) >= 14
    seen_pipes, seen_shard, seen_kernel = set(), set(), set()
    for f in files:
        cfg = B._merge(B.DEFAULT, yaml.safe_load(open(f)))
        seen_pipes.update(cfg["run"]["pipelines"])
        seen_shard.add(cfg["run"]["sharding"])
        seen_kernel.add(cfg["run"]["compute_kernel"])
        assert all(p in B.PIPELINES for p in cfg["run"]["pipelines"]), f
        small = B._merge(cfg, {"model": {"dense_arch": [16], "over_arch": [16, 1], "embedding_dim": 8}, "tables": {"num": 3, "rows": 50, "pooling": 2},
                               "run": {"batch_size": 8, "steps": 2, "warmup": 1, "dense_backend": "torch", "local_world_size": 0, "grad_accumulation": min(2, cfg["run"]["grad_accumulation"])}})
        res = B.run(small)
        assert len(res) == len(small["run"]["pipelines"]) and all(r["ms_per_step"] > 0 for r in res), f
    assert {"base", "sparse_dist", "sparse_dist_lite", "fused_sparse_dist", "semi_sync", "prefetch", "emb_stash", "opt_stash", "bwd_opt"} <= seen_pipes
    assert {"table_wise", "row_wise", "column_wise", "table_row_wise", "planner"} <= seen_shard and {"fused", "fused_uvm_caching", "key_value"} <= seen_kernel


def test_pt2_utils_transformer_compile_switch_and_queue():
    import torch

    from torchrec_b200.pt2.utils import AtomicCounter, TensorQueue, default_pipeline_input_transformer, pt2_compile_callable, register_fake_classes
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor

    class B:
        pass

    b = B()
    b.id_list_features = KeyedJaggedTensor(keys=["f"], values=torch.tensor([1, 2, 3]), lengths=torch.tensor([2, 1], dtype=torch.int32))
    b.other = 5
    out = default_pipeline_input_transformer(b)
    assert out is b and out.id_list_features.lengths().dtype == torch.int64 and out.other == 5

    calls = []

    class M:
        enable_pt2_compile = False

        @pt2_compile_callable
        def update(self, x):
            calls.append("eager")
            return x * 2

    m = M()
    assert torch.equal(m.update(torch.ones(2)), torch.full((2,), 2.0)) and calls == ["eager"] and "_update_pt2_compiled" not in m.__dict__
    q = TensorQueue(torch.zeros(1))
    assert q.size() == 0 and torch.equal(q.pop(), torch.zeros(1))
    q.push(torch.ones(1))
    q.push(torch.full((1,), 2.0))
    assert torch.equal(q.top(), torch.ones(1)) and torch.equal(q.pop(), torch.ones(1)) and q.size() == 1
    c = AtomicCounter()
    assert c.increment() == 1 and c.increment() == 2 and c.decrement() == 1
    c.set(7)
    assert c.get() == 7
    register_fake_classes()


def test_distributed_benchmark_package(tmp_path):
    """``torchrec_b200.distributed.benchmark``: the harness modules under the reference's names, the micro-benchmarks run on CPU, trace /
    snapshot post-processing on synthetic files."""
    import importlib
    import json
    import pickle

    import torch

    base = importlib.import_module("torchrec_b200.distributed.benchmark.base")
    import torchrec_b200.benchmarks.base as real_base

    assert base is real_base and importlib.import_module("torchrec_b200.distributed.benchmark.benchmark_train_pipeline").__name__ == "torchrec_b200.benchmarks.benchmark_train_pipeline"
    from torchrec_b200.distributed.benchmark import utils as bu
    from torchrec_b200.distributed.benchmark.benchmark_set_sharding_context_post_a2a import _set_sharding_context_post_a2a_previous, op_bench as ctx_bench
    from torchrec_b200.distributed.benchmark.benchmark_split_table_batched_embeddings import op_bench as tbe_bench
    from torchrec_b200.distributed.benchmark.benchmark_train import benchmark_ec_write
    from torchrec_b200.distributed.benchmark.embedding_collection_wrappers import benchmark_ebc_module, get_tables
    from torchrec_b200.parallel.embedding_sharding import _set_sharding_context_post_a2a

    r = tbe_bench(1000, 16, 2, 32, 4, num_benchmarks=2, device="cpu")
    assert r.gpu_elapsed_time.numel() == 2 and r.qps is not None and r.cpu_mem_stats and "fwdbwd" in r.short_name
    assert ctx_bench(20, 5, _set_sharding_context_post_a2a_previous)["ms"] > 0 and ctx_bench(20, 5, _set_sharding_context_post_a2a)["ms"] > 0
    w = benchmark_ec_write(num_embeddings=500, embedding_dim=8, num_tables=2, batch_size=16, iters=2, device=torch.device("cpu"))
    assert "ec_write_read" in w.short_name
    # a sharded EBC under two sharding types, 2 gloo ranks
    from torchrec_b200.modules.embedding_configs import DataType
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.types import ShardingType

    tables = get_tables([(200, 8), (100, 8)], data_type=DataType.FP32)
    ebc = EmbeddingBagCollection(tables)
    apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": 0.1})
    res = benchmark_ebc_module(ebc, EmbeddingBagCollectionSharder(), [ShardingType.TABLE_WISE, ShardingType.ROW_WISE], [base.CompileMode.EAGER], tables, warmup_iters=1,
                               bench_iters=2, prof_iters=1, batch_size=8, world_size=2, num_benchmarks=2, device_type="cpu")
    assert [x.short_name for x in res] == ["table_wise-eager", "row_wise-eager"] and all(x.gpu_elapsed_time.numel() == 2 for x in res)
    report = tmp_path / "report.txt"
    base.write_report(res, str(report), "header\n", num_requests=16)
    assert "table_wise-eager" in report.read_text() and "Avg QPS" in report.read_text()
    out = bu.dump_benchmark_result(res[0], str(tmp_path / "json"), extra={"note": "x"})
    assert json.load(open(out))["name"] == "table_wise-eager"
    # chrome trace: two streams, [0, 10) + [5, 20) on stream 7, [30, 40) on stream 9 over a span of 40 us
    trace = {"traceEvents": [{"ph": "X", "cat": "kernel", "ts": 0, "dur": 10, "args": {"stream": 7}}, {"ph": "X", "cat": "kernel", "ts": 5, "dur": 15, "args": {"stream": 7}},
                             {"ph": "X", "cat": "gpu_memcpy", "ts": 30, "dur": 10, "args": {"stream": 9}}, {"ph": "X", "cat": "cpu_op", "ts": 0, "dur": 100}]}
    tp = tmp_path / bu.create_trace_file_name("p", 0)
    tp.write_text(json.dumps(trace))
    u = bu.parse_chrome_trace_gpu_utilization(str(tp))
    assert u["gpu_utilization"] == 30 / 40 and u["stream_7_utilization"] == 20 / 40 and u["stream_9_utilization"] == 10 / 40
    snap = {"device_traces": [[{"action": "alloc", "size": 4 << 20, "stream": 0}, {"action": "alloc", "size": 2 << 20, "stream": 3}, {"action": "free_completed", "size": 4 << 20, "stream": 0},
                               {"action": "alloc", "size": 1 << 20, "stream": 0}]]}
    sp_ = tmp_path / bu.create_snapshot_file_name("p", 0)
    sp_.write_bytes(pickle.dumps(snap))
    assert bu.parse_memory_snapshot_peak_per_stream(str(sp_)) == {"stream_0_peak_mb": 4.0, "stream_3_peak_mb": 2.0, "total_peak_mb": 6.0}
    assert bu.get_cpu_type() and bu.get_gpu_type()
