"""Benchmark post-processing: device names, trace / snapshot file names, GPU utilisation out of a chrome trace, peak memory per stream
out of an allocator snapshot, results as json (reference ``distributed/benchmark/utils.py``)."""
from __future__ import annotations

import gzip
import json
import os
import pickle
import platform
from typing import Any, Dict, List, Optional, Tuple

import torch


def get_gpu_type() -> str:
    return torch.cuda.get_device_name(0) if torch.cuda.is_available() else "cpu"


def get_cpu_type() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or platform.machine()


def create_trace_file_name(profile_name: str, rank: int) -> str:
    return f"trace-{profile_name}-rank{rank}.json"


def create_snapshot_file_name(profile_name: str, rank: int) -> str:
    return f"memory-{profile_name}-rank{rank}.pickle"


def _load_trace_events(trace_path: str) -> List[Dict[str, Any]]:
    opener = gzip.open if trace_path.endswith(".gz") else open
    with opener(trace_path, "rt") as f:
        data = json.load(f)
    return data["traceEvents"] if isinstance(data, dict) else data


def _extract_stream_tracks(events: List[Dict[str, Any]]) -> Dict[Any, List[Tuple[float, float]]]:
    """Complete events of device activity (kernels, memcpy / memset) grouped by the stream they ran on: (start, end) in microseconds."""
    tracks: Dict[Any, List[Tuple[float, float]]] = {}
    for e in events:
        if e.get("ph") != "X" or e.get("cat") not in ("kernel", "gpu_memcpy", "gpu_memset", "Kernel", "Memcpy", "Memset"):
            continue
        stream = (e.get("args") or {}).get("stream", e.get("tid"))
        tracks.setdefault(stream, []).append((float(e["ts"]), float(e["ts"]) + float(e.get("dur", 0.0))))
    return tracks


def _merged_active_time(intervals: List[Tuple[float, float]]) -> float:
    """Length of the union of the intervals."""
    total, cur_s, cur_e = 0.0, None, None
    for s, e in sorted(intervals):
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                total += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        total += cur_e - cur_s
    return total


def parse_chrome_trace_gpu_utilization(trace_path: str) -> Dict[str, float]:
    """Share of the traced span during which the device ran something (any stream), and per stream: ``gpu_utilization`` and
    ``stream_<id>_utilization`` in [0, 1]; ``gpu_active_us`` / ``span_us`` for reference."""
    tracks = _extract_stream_tracks(_load_trace_events(trace_path))
    all_iv = [iv for ivs in tracks.values() for iv in ivs]
    if not all_iv:
        return {"gpu_utilization": 0.0, "gpu_active_us": 0.0, "span_us": 0.0}
    span = max(e for _, e in all_iv) - min(s for s, _ in all_iv)
    out = {"gpu_utilization": _merged_active_time(all_iv) / span if span > 0 else 0.0, "gpu_active_us": _merged_active_time(all_iv), "span_us": span}
    for stream, ivs in tracks.items():
        out[f"stream_{stream}_utilization"] = _merged_active_time(ivs) / span if span > 0 else 0.0
    return out


def _merge_gpu_utilization_metrics(per_rank: List[Dict[str, float]]) -> Dict[str, float]:
    keys = set().union(*[m.keys() for m in per_rank]) if per_rank else set()
    return {k: sum(m.get(k, 0.0) for m in per_rank) / len(per_rank) for k in keys}


def parse_memory_snapshot_peak_per_stream(snapshot_path: str) -> Dict[str, float]:
    """Peak of the live allocated bytes per stream, replayed from the alloc / free events of a ``torch.cuda.memory._dump_snapshot``
    pickle: ``{"stream_<id>_peak_mb": ..., "total_peak_mb": ...}``."""
    with open(snapshot_path, "rb") as f:
        snap = pickle.load(f)
    live: Dict[Any, int] = {}
    peak: Dict[Any, int] = {}
    total = total_peak = 0
    for trace in snap.get("device_traces", []):
        for ev in trace:
            action, size, stream = ev.get("action"), int(ev.get("size", 0)), ev.get("stream", 0)
            if action == "alloc":
                live[stream] = live.get(stream, 0) + size
                total += size
            elif action in ("free_completed", "free"):
                live[stream] = live.get(stream, 0) - size
                total -= size
            else:
                continue
            peak[stream] = max(peak.get(stream, 0), live[stream])
            total_peak = max(total_peak, total)
    out = {f"stream_{s}_peak_mb": p / 2**20 for s, p in peak.items()}
    out["total_peak_mb"] = total_peak / 2**20
    return out


def dump_benchmark_result(result: Any, output_dir: str, extra: Optional[Dict[str, Any]] = None) -> str:
    """``<output_dir>/<short_name>.json``: percentiles of device / host time, memory per rank, device names."""
    os.makedirs(output_dir, exist_ok=True)
    doc = {"name": result.short_name, "rank": result.rank, "gpu_type": get_gpu_type(), "cpu_type": get_cpu_type(),
           "gpu_ms": {f"p{p}": result.runtime_percentile(p) for p in (50, 75, 90, 99)} if result.gpu_elapsed_time.numel() else {},
           "cpu_ms": {f"p{p}": result.runtime_percentile(p, "cpu") for p in (50, 75, 90, 99)} if result.cpu_elapsed_time.numel() else {},
           "gpu_mem": [str(m) if not isinstance(m, dict) else {"allocated_peak": m.get("allocated_bytes.all.peak", 0)} for m in result.gpu_mem_stats],
           "cpu_mem": [str(m) for m in getattr(result, "cpu_mem_stats", [])], "qps": getattr(result, "qps", None)}
    if extra:
        doc.update(extra)
    path = os.path.join(output_dir, f"{result.short_name}.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    return path
