"""Benchmark harness: dataclass-driven CLI (``cmd_conf``), CUDA-event timing with CPU-active time and RSS, optional
``torch.profiler`` chrome traces and CUDA memory-history snapshots.

Parity: reference ``distributed/benchmark/base.py`` — ``cmd_conf`` (:563-760: dataclass fields become ``--flags``, values can come
from ``--yaml_config`` / ``--json_config`` files, CLI wins), ``BenchmarkResult`` (:100-260), the ``benchmark_func`` timing loop with
profiler + memory snapshot (:788-1090). Rebuilt around CUDA events on the launching stream (device time, max over ranks is the
caller's job) instead of wall clock."""
from __future__ import annotations

import argparse
import dataclasses
import functools
import inspect
import json
import os
import resource
import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Union, get_args, get_origin, get_type_hints

import torch


@dataclass
class BenchmarkResult:
    short_name: str
    gpu_elapsed_time: torch.Tensor  # ms per iteration (device, CUDA events)
    cpu_elapsed_time: torch.Tensor  # ms per iteration the host spent enqueueing
    gpu_mem_stats: List[Dict[str, int]] = field(default_factory=list)
    rank: int = -1
    cpu_rss_mb: float = 0.0

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
