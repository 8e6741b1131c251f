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
