"""Benchmark harness: dataclass-driven CLI (``cmd_conf``), CUDA-event timing with CPU-active time and RSS, optional
``torch.profiler`` chrome traces and CUDA memory-history snapshots.

Parity: reference ``distributed/benchmark/base.py`` — ``cmd_conf`` (:563-760: dataclass fields become ``--flags``, values can come
from ``--yaml_config`` / ``--json_config`` files, CLI wins), ``BenchmarkResult`` (:100-260), the ``benchmark_func`` timing loop with
profiler + memory snapshot (:788-1090). Rebuilt around CUDA events on the launching stream (device time, max over ranks is the
caller's job) instead of wall clock."""
from __future__ import annotations

import argparse
import dataclasses
import logging
import multiprocessing as mp
from enum import Enum
import functools
import inspect
import json
import os
import resource
import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple, Union, get_args, get_origin, get_type_hints

import torch


class CompileMode(Enum):
    """How a benchmarked module is prepared. This framework runs eager modules on hand-written kernels and CUDA graphs; ``FX_SCRIPT``
    (fx trace, then TorchScript where the module scripts) is kept for inference modules."""

    EAGER = "eager"
    FX_SCRIPT = "fx_script"


@dataclass
class GPUMemoryStats:
    rank: int
    malloc_retries: int
    max_mem_allocated_mbs: int
    max_mem_reserved_mbs: int
    free_mbs: int
    total_mbs: int

    @classmethod
    def for_device(cls, rank: int) -> "GPUMemoryStats":
        stats = torch.cuda.memory_stats(rank)
        free, total = torch.cuda.mem_get_info(rank)
        mb = 1024 * 1024
        return cls(rank, stats.get("num_alloc_retries", 0), stats.get("allocated_bytes.all.peak", 0) // mb, stats.get("reserved_bytes.all.peak", 0) // mb, free // mb, total // mb)

    def get(self, key: str, default: int = 0) -> int:  # dict-style access of the raw allocator statistic this summarises
        return {"allocated_bytes.all.peak": self.max_mem_allocated_mbs * 1024 * 1024, "reserved_bytes.all.peak": self.max_mem_reserved_mbs * 1024 * 1024,
                "num_alloc_retries": self.malloc_retries}.get(key, default)

    def __str__(self) -> str:
        used = self.total_mbs - self.free_mbs
        return (f"GPUMemoryStats: Rank {self.rank}: retries={self.malloc_retries}, allocated={self.max_mem_allocated_mbs:6}mb, reserved={self.max_mem_reserved_mbs:6}mb, "
                f"free={self.free_mbs:6}mb, total={self.total_mbs:6}mb, used={used:6}mb overhead={used - self.max_mem_reserved_mbs:6}mb")


@dataclass
class CPUMemoryStats:
    rank: int
    peak_rss_mbs: int

    @classmethod
    def for_process(cls, rank: int) -> "CPUMemoryStats":
        return cls(rank, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss // 1024)

    def __str__(self) -> str:
        return f"Rank {self.rank}: CPU Memory Peak RSS: {self.peak_rss_mbs / 1000:.2f} GB"


@dataclass
class BenchmarkResult:
    short_name: str
    gpu_elapsed_time: torch.Tensor  # ms per iteration (device, CUDA events)
    cpu_elapsed_time: torch.Tensor  # ms per iteration the host spent enqueueing
    gpu_mem_stats: List[Any] = field(default_factory=list)  # raw allocator statistics (dict) or GPUMemoryStats, one per rank
    rank: int = -1
    cpu_rss_mb: float = 0.0
    cpu_mem_stats: List[CPUMemoryStats] = field(default_factory=list)
    qps: Optional[float] = None
    cpu_utilization: Optional[float] = None  # host time enqueueing / device time

    def runtime_percentile(self, percentile: int = 50, device: str = "gpu") -> float:
        t = self.gpu_elapsed_time if device == "gpu" else self.cpu_elapsed_time
        return float(torch.quantile(t.double(), percentile / 100.0))

    def max_mem_alloc_by_rank(self) -> List[int]:
        return [s.get("allocated_bytes.all.peak", 0) for s in self.gpu_mem_stats]

    def __str__(self) -> str:
        mem = f" | peak HBM {max(self.max_mem_alloc_by_rank()) / 2**30:.2f} GiB" if self.gpu_mem_stats else ""
        return (f"{self.short_name: <28} | GPU p50 {self.runtime_percentile(50):8.3f} ms p90 {self.runtime_percentile(90):8.3f} ms"
                f" | CPU p50 {self.runtime_percentile(50, 'cpu'):8.3f} ms{mem} | RSS {self.cpu_rss_mb:.0f} MB")


def benchmark_func(name: str, fn: Callable[[], Any], num_benchmarks: int = 10, num_warmup: int = 3, device: Optional[torch.device] = None,
                   profile_dir: str = "", memory_snapshot: bool = False, rank: int = 0) -> BenchmarkResult:
    """Time ``fn`` (one iteration per call). With ``profile_dir``: a chrome trace of 3 extra iterations is written there; with
    ``memory_snapshot``: the CUDA caching-allocator history of those iterations is dumped as a pickle for
    pytorch.org/memory_viz."""
    cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
    for _ in range(num_warmup):
        fn()
    if cuda:
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
    gpu_ms, cpu_ms = [], []
    for _ in range(num_benchmarks):
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        fn()
        cpu_ms.append((time.perf_counter() - t0) * 1e3)
        if cuda:
            e1.record()
            e1.synchronize()
            gpu_ms.append(e0.elapsed_time(e1))
        else:
            gpu_ms.append(cpu_ms[-1])
    stats = [dict(torch.cuda.memory_stats())] if cuda else []
    if profile_dir:
        os.makedirs(profile_dir, exist_ok=True)
        from torch.profiler import ProfilerActivity, profile

        if cuda and memory_snapshot:
            torch.cuda.memory._record_memory_history(max_entries=100000)
        acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if cuda else [])
        with profile(activities=acts, record_shapes=False) as prof:
            for _ in range(3):
                fn()
            if cuda:
                torch.cuda.synchronize()
        prof.export_chrome_trace(os.path.join(profile_dir, f"trace-{name}-rank{rank}.json"))
        if cuda and memory_snapshot:
            torch.cuda.memory._dump_snapshot(os.path.join(profile_dir, f"memory-{name}-rank{rank}.pickle"))
            torch.cuda.memory._record_memory_history(enabled=None)
    rss = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0
    return BenchmarkResult(short_name=name, gpu_elapsed_time=torch.tensor(gpu_ms), cpu_elapsed_time=torch.tensor(cpu_ms), gpu_mem_stats=stats, rank=rank, cpu_rss_mb=rss)


# ---- dataclass -> argparse (+ YAML / JSON defaults) --------------------------------------------------------------------------
def _unwrap_optional(tp: Any) -> Any:
    if get_origin(tp) is Union:
        args = [a for a in get_args(tp) if a is not type(None)]
        if len(args) == 1:
            return args[0]
    return tp


def _add_dataclass_args(parser: argparse.ArgumentParser, cls: type, seen: Dict[str, type]) -> None:
    hints = get_type_hints(cls)
    for f in dataclasses.fields(cls):
        if f.name in seen:
            continue
        seen[f.name] = cls
        tp = _unwrap_optional(hints.get(f.name, str))
        default = None  # real defaults are applied after the config file is merged
        if tp is bool:
            parser.add_argument(f"--{f.name}", type=lambda s: str(s).lower() in ("1", "true", "yes", "y"), default=default, metavar="BOOL")
        elif get_origin(tp) in (list, List):
            inner = (get_args(tp) or (str,))[0]
            parser.add_argument(f"--{f.name}", type=inner if inner in (int, float, str) else str, nargs="*", default=default)
        elif isinstance(tp, type) and issubclass(tp, (int, float, str)):
            parser.add_argument(f"--{f.name}", type=tp, default=default)
        else:
            parser.add_argument(f"--{f.name}", type=str, default=default, help="JSON value")


def _load_config_file(path: str) -> Dict[str, Any]:
    with open(path) as f:
        if path.endswith((".yml", ".yaml")):
            import yaml

            return yaml.safe_load(f) or {}
        return json.load(f)


def cmd_conf(func: Callable[..., Any]) -> Callable[..., Any]:
    """Turn a function whose parameters are dataclasses into a CLI entry point: every dataclass field becomes ``--field``;
    ``--yaml_config`` / ``--json_config`` files provide defaults (flat ``field: value`` pairs or ``{DataclassName: {...}}``
    sections); precedence CLI > file > dataclass default. Non-dataclass parameters become plain flags."""
    sig = inspect.signature(func)
    hints = get_type_hints(func)

    @functools.wraps(func)
    def wrapper(argv: Optional[List[str]] = None) -> Any:
        parser = argparse.ArgumentParser(description=func.__doc__)
        parser.add_argument("--yaml_config", type=str, default=None)
        parser.add_argument("--json_config", type=str, default=None)
        seen: Dict[str, type] = {}
        plain: Dict[str, inspect.Parameter] = {}
        for name, prm in sig.parameters.items():
            tp = hints.get(name)
            if tp is not None and dataclasses.is_dataclass(tp):
                _add_dataclass_args(parser, tp, seen)
            else:
                plain[name] = prm
                parser.add_argument(f"--{name}", type=tp if tp in (int, float, str) else str, default=None)
        ns, _unknown = parser.parse_known_args(argv)
        file_cfg: Dict[str, Any] = {}
        for path in (ns.yaml_config, ns.json_config):
            if path:
                file_cfg.update(_load_config_file(path))
        kwargs: Dict[str, Any] = {}
        for name, prm in sig.parameters.items():
            tp = hints.get(name)
            if tp is not None and dataclasses.is_dataclass(tp):
                section = file_cfg.get(tp.__name__, {}) if isinstance(file_cfg.get(tp.__name__), dict) else {}
                vals: Dict[str, Any] = {}
                for f in dataclasses.fields(tp):
                    cli = getattr(ns, f.name, None)
                    if cli is not None:
                        vals[f.name] = cli
                    elif f.name in section:
                        vals[f.name] = section[f.name]
                    elif f.name in file_cfg and not isinstance(file_cfg[f.name], dict):
                        vals[f.name] = file_cfg[f.name]
                kwargs[name] = tp(**vals)
            else:
                cli = getattr(ns, name, None)
                if cli is not None:
                    kwargs[name] = cli
                elif name in file_cfg:
                    kwargs[name] = file_cfg[name]
                elif prm.default is not inspect.Parameter.empty:
                    kwargs[name] = prm.default
        return func(**kwargs)

    return wrapper


# ---- reports, multi-process runs, reference-shaped entry points ---------------------------------------------------------------------------------------
logger = logging.getLogger(__name__)
EMBEDDING_DIM = 128
DLRM_NUM_EMBEDDINGS_PER_FEATURE = [4833188, 36746, 17245, 7413, 20243, 3, 7114, 1441, 62, 29275261, 1572176, 345138, 10, 2209, 11267, 128, 4, 974, 14, 48937457,
                                   11316796, 40094537, 452104, 12606, 104, 35]


def write_report(benchmark_results: List[BenchmarkResult], report_file: str, report_str: str, num_requests: int) -> None:
    """One line per result (QPS over the device time, device / host mean +- std) + the memory statistics of every rank."""
    for res in benchmark_results:
        g, c = res.gpu_elapsed_time.double() * 1e-3, res.cpu_elapsed_time.double() * 1e-3
        std = lambda t: float(t.std()) if t.numel() > 1 else 0.0  # noqa: E731
        mem = "".join(f"{m}\n" for m in list(res.gpu_mem_stats) + list(res.cpu_mem_stats))
        report_str += (f"{res.short_name:40} Avg QPS(GPU):{int(num_requests / max(float(g.mean()), 1e-12)):10} GPU Avg: {1000 * float(g.mean()):8.2f}ms ±{1000 * std(g):.2f}ms "
                       f"CPU Avg: {1000 * float(c.mean()):8.2f}ms ±{1000 * std(c):.2f}ms\n\tMemory Allocated Per Rank:\n\t{mem}\n")
    with open(report_file, "w") as f:
        f.write(report_str)
    logger.info("Report written to %s:\n%s", report_file, report_str)


def multi_process_benchmark(callable: Callable[..., None], **kwargs: Any) -> BenchmarkResult:
    """Run ``callable(rank=r, world_size=W, queue=q, **kwargs)`` in W spawned processes (rendezvous on 127.0.0.1); every rank puts its
    ``BenchmarkResult`` on the queue; returned: rank 0's timings with the memory statistics of all ranks."""
    from ..utils.multiprocess import get_free_port as free_port

    assert "world_size" in kwargs
    world_size = kwargs["world_size"]
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(free_port()))
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = []
    for rank in range(world_size):
        p = ctx.Process(target=callable, kwargs={**kwargs, "rank": rank, "world_size": world_size, "queue": q})
        p.start()
        procs.append(p)
    per_rank = [q.get() for _ in range(world_size)]
    for p in procs:
        p.join()
        assert p.exitcode == 0, f"benchmark process exited with {p.exitcode}"
    per_rank.sort(key=lambda r: r.rank)
    first = per_rank[0]
    return BenchmarkResult(short_name=first.short_name, gpu_elapsed_time=first.gpu_elapsed_time, cpu_elapsed_time=first.cpu_elapsed_time, rank=0, qps=first.qps,
                           cpu_utilization=first.cpu_utilization, cpu_rss_mb=first.cpu_rss_mb, gpu_mem_stats=[m for r in per_rank for m in r.gpu_mem_stats[:1]],
                           cpu_mem_stats=[m for r in per_rank for m in (r.cpu_mem_stats[:1] or [CPUMemoryStats(r.rank, int(r.cpu_rss_mb))])])


def set_embedding_config(embedding_config_json: str) -> Tuple[List[Tuple[int, int]], List[int]]:
    """``{feature: {num_embeddings, embedding_dim[, pooling_factor]}}`` from a json file -> ([(rows, dim)], [pooling factors]); the
    Criteo-1TB DLRM tables at dim 128 when the file is missing or malformed."""
    configs: List[Tuple[int, int]] = []
    pooling: List[int] = []
    try:
        with open(embedding_config_json) as f:
            data = json.load(f)
        for cfg in data.values():
            configs.append((cfg["num_embeddings"], cfg["embedding_dim"]))
            if "pooling_factor" in cfg:
                pooling.append(cfg["pooling_factor"])
            elif pooling:
                raise RuntimeError("We cannot handle some features have pooling factor and others don't.")
        if pooling and len(pooling) != len(configs):
            raise RuntimeError("We cannot handle some features have pooling factor and others don't.")
    except BaseException as e:  # noqa: B036
        logger.warning("Failed to load embedding config because %s, fallback to DLRM config", e)
        configs, pooling = [(n, EMBEDDING_DIM) for n in DLRM_NUM_EMBEDDINGS_PER_FEATURE], []
    return configs, pooling


@dataclass
class BenchFuncConfig:
    """The knobs of one ``benchmark_func`` run, as a config-file / command-line section."""

    name: str
    world_size: int
    num_profiles: int
    num_benchmarks: int
    profile_dir: str = ""
    device_type: str = "cuda"
    pre_gpu_load: int = 0
    export_stacks: bool = False
    all_rank_traces: bool = False
    memory_snapshot: bool = False
    loglevel: str = "WARNING"

    def benchmark_func_kwargs(self, **kwargs_to_override: Any) -> Dict[str, Any]:
        return {"name": self.name, "world_size": self.world_size, "num_profiles": self.num_profiles, "num_benchmarks": self.num_benchmarks, "profile_dir": self.profile_dir,
                "device_type": self.device_type, "pre_gpu_load": self.pre_gpu_load, "export_stacks": self.export_stacks, "all_rank_traces": self.all_rank_traces,
                "memory_snapshot": self.memory_snapshot} | kwargs_to_override

    def set_log_level(self) -> None:
        logging.root.setLevel(logging._nameToLevel[self.loglevel.upper()])


def _pre_gpu_load(pre_gpu_load: int, device_type: str) -> None:
    """Dummy matmuls before the first measured iteration (a busy allocator / warm clocks)."""
    if pre_gpu_load and device_type == "cuda" and torch.cuda.is_available():
        x = torch.rand(16384, 16384, device="cuda")
        for _ in range(pre_gpu_load):
            x = x * torch.rand(16384, 16384, device="cuda")


def benchmark_inputs_func(name: str, rank: int, world_size: int, func_to_benchmark: Any, bench_inputs: List[Any], prof_inputs: List[Any],
                          benchmark_func_kwargs: Optional[Dict[str, Any]], num_profiles: int, num_benchmarks: int, profile_dir: str, device_type: str = "cuda",
                          pre_gpu_load: int = 0, export_stacks: bool = False, all_rank_traces: bool = False, memory_snapshot: bool = False, sample_count: int = 0) -> BenchmarkResult:
    """The reference's ``benchmark_func`` calling convention: one measured iteration = ``func_to_benchmark(bench_inputs, **kwargs)``;
    profiling (rank 0, or all ranks) runs it on ``prof_inputs``. ``sample_count`` > 0 adds examples / second to the result."""
    kw = benchmark_func_kwargs or {}
    _pre_gpu_load(pre_gpu_load, device_type)
    res = benchmark_func(name, lambda: func_to_benchmark(bench_inputs, **kw), num_benchmarks=num_benchmarks, num_warmup=1, device=torch.device(device_type),
                         profile_dir="", rank=max(rank, 0))
    if profile_dir and (all_rank_traces or rank <= 0):
        benchmark_func(name, lambda: func_to_benchmark(prof_inputs, **kw), num_benchmarks=0, num_warmup=0, device=torch.device(device_type), profile_dir=profile_dir,
                       memory_snapshot=memory_snapshot, rank=max(rank, 0))
    _finish(res, rank, device_type, sample_count)
    return res


def _finish(res: BenchmarkResult, rank: int, device_type: str, sample_count: int) -> None:
    res.cpu_mem_stats = [CPUMemoryStats.for_process(max(rank, 0))]
    if device_type == "cuda" and torch.cuda.is_available():
        res.gpu_mem_stats = [GPUMemoryStats.for_device(torch.cuda.current_device())]
    g = float(res.gpu_elapsed_time.double().mean()) if res.gpu_elapsed_time.numel() else 0.0
    res.cpu_utilization = float(res.cpu_elapsed_time.double().mean()) / g if g > 0 else None
    if sample_count and g > 0:
        res.qps = sample_count / (g * 1e-3)


def benchmark_model_with_warmup(name: str, model: torch.nn.Module, warmup_inputs: List[Any], bench_inputs: List[Any], prof_inputs: List[Any], world_size: int, output_dir: str,
                                num_benchmarks: int, func_to_benchmark: Any, benchmark_func_kwargs: Optional[Dict[str, Any]], rank: int, enable_logging: bool = True,
                                device_type: str = "cuda", benchmark_unsharded_module: bool = False, export_stacks: bool = False) -> BenchmarkResult:
    """Warm the model on ``warmup_inputs``, then time ``func_to_benchmark(model, bench_inputs, **kwargs)``; with ``output_dir`` a chrome
    trace of a pass over ``prof_inputs`` is written there."""
    if enable_logging:
        logger.info(" BENCHMARK_MODEL[%s]:\n%s", name, model)
    for x in warmup_inputs:
        model(x)
    kw = benchmark_func_kwargs or {}
    res = benchmark_func(name, lambda: func_to_benchmark(model, bench_inputs, **kw), num_benchmarks=num_benchmarks, num_warmup=0, device=torch.device(device_type), rank=max(rank, 0))
    if output_dir:
        benchmark_func(name, lambda: [model(x) for x in prof_inputs], num_benchmarks=0, num_warmup=0, device=torch.device(device_type), profile_dir=output_dir, rank=max(rank, 0))
    _finish(res, rank, device_type, 0)
    return res


def init_argparse_and_args(argv: Optional[List[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser()
    parser.add_argument("--warmup_iters", type=int, default=20)
    parser.add_argument("--bench_iters", type=int, default=500)
    parser.add_argument("--prof_iters", type=int, default=20)
    parser.add_argument("--batch_size", type=int, default=2048)
    parser.add_argument("--world_size", type=int, default=2)
    parser.add_argument("--max_num_embeddings", type=int, default=1000000)
    parser.add_argument("--output_dir", type=str, default="/var/tmp/torchrec-bench")
    parser.add_argument("--num_benchmarks", type=int, default=5)
    parser.add_argument("--embedding_config_json", type=str, default="")
    parser.add_argument("--device_type", type=str, default="cuda")
    return parser.parse_args(argv)
