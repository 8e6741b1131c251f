"""Sharded checkpoint writer / loader with re-sharding on load.

The reference ships no writer (users hand the ShardedTensor state dicts to torch.distributed.checkpoint).
This module is the built-in equivalent for one-process-per-GPU jobs: every rank writes ONE file with its local shards
(weights + fused optimizer state, any dtype) and rank 0 writes the manifest (global shapes, shard rectangles, dense
tensors). Loading works under ANY new sharding plan or world size: each rank reads exactly the byte ranges that intersect
its new shards (rectangle intersection, the same primitive as dynamic re-sharding).

    save(model, fused_optimizer, "ckpt_dir")            # collective
    load(model, fused_optimizer, "ckpt_dir")            # collective, plan / world size may differ
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

_MANIFEST = "manifest.json"


def _rank_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _is_sharded(t: Any) -> bool:
    return hasattr(t, "local_shards") and hasattr(t, "metadata")


def _shards_of(t: Any) -> List[Tuple[torch.Tensor, List[int], List[int]]]:
    return [(s.tensor, list(s.metadata.shard_offsets), list(s.metadata.shard_sizes)) for s in t.local_shards()]


def _global_size(t: Any) -> List[int]:
    return list(t.metadata().size) if callable(getattr(t, "metadata", None)) else list(t.size())


def _flatten_optim(sd: Dict[str, Any]) -> Dict[str, Any]:
    flat: Dict[str, Any] = {}
    for pname, states in sd.get("state", {}).items():
        for sname, v in states.items():
            if isinstance(v, torch.Tensor) or _is_sharded(v):
                flat[f"{pname}::{sname}"] = v
    return flat


def save(model: torch.nn.Module, optimizer: Optional[Any], path: str, extra: Optional[Dict[str, Any]] = None) -> None:
    rank, world = _rank_world()
    os.makedirs(path, exist_ok=True)
    entries: Dict[str, Any] = dict(model.state_dict())
    if optimizer is not None:
        entries.update({f"__optim__/{k}": v for k, v in _flatten_optim(optimizer.state_dict()).items()})
    blob: Dict[str, torch.Tensor] = {}
    local_index: Dict[str, List[Dict[str, Any]]] = {}
    dense: Dict[str, torch.Tensor] = {}
    for key, v in entries.items():
        if _is_sharded(v):
            for i, (t, off, size) in enumerate(_shards_of(v)):
                name = f"{key}@@{i}"
                blob[name] = t.detach().cpu().contiguous()
                local_index.setdefault(key, []).append({"blob": name, "offsets": off, "sizes": size, "rank": rank, "global": _global_size(v)})
        elif isinstance(v, torch.Tensor) and rank == 0:
            dense[key] = v.detach().cpu()
    torch.save(blob, os.path.join(path, f"shards_rank{rank}.pt"))
    gathered: List[Any] = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, local_index)
    else:
        gathered = [local_index]
    if rank == 0:
        index: Dict[str, List[Dict[str, Any]]] = {}
        for part in gathered:
            for k, lst in part.items():
                index.setdefault(k, []).extend(lst)
        torch.save(dense, os.path.join(path, "dense.pt"))
        with open(os.path.join(path, _MANIFEST), "w") as f:
            json.dump({"world_size": world, "sharded": index, "dense": sorted(dense.keys()), "extra": extra or {}}, f)
    if world > 1:
        dist.barrier()


def _fill(dst: torch.Tensor, dst_off: List[int], dst_size: List[int], pieces: List[Dict[str, Any]], blobs: Dict[int, Dict[str, torch.Tensor]], path: str,
          dst_global: Optional[List[int]] = None) -> int:
    """Copy every saved rectangle intersecting [dst_off, dst_off+dst_size) into dst. Returns covered elements.

    Row-wise optimizer state of column-sharded tables is a 1-D tensor of ``rows x n_column_shards`` values (one copy of the
    per-row state per column shard, reference batched_embedding_kernel.py:1259-1329). When the saved and the target layout
    disagree on that multiplicity, both are folded back onto the table's row axis (column shard 0 is the representative)."""
    covered = 0
    fold = None
    if len(dst_off) == 1 and dst_global is not None and pieces and pieces[0]["global"][0] != dst_global[0]:
        fold = min(pieces[0]["global"][0], dst_global[0])
        dst_off = [dst_off[0] % fold]
    for p in pieces:
        so, ss = p["offsets"], p["sizes"]
        if fold is not None:
            if so[0] // fold != 0:
                continue
            so = [so[0] % fold]
        lo = [max(a, b) for a, b in zip(dst_off, so)]
        hi = [min(a + x, b + y) for a, x, b, y in zip(dst_off, dst_size, so, ss)]
        if any(l >= h for l, h in zip(lo, hi)):
            continue
        r = p["rank"]
        if r not in blobs:
            blobs[r] = torch.load(os.path.join(path, f"shards_rank{r}.pt"), map_location="cpu", mmap=True)
        src = blobs[r][p["blob"]]
        s_idx = tuple(slice(l - o, h - o) for l, h, o in zip(lo, hi, so))
        d_idx = tuple(slice(l - o, h - o) for l, h, o in zip(lo, hi, dst_off))
        dst[d_idx].copy_(src[s_idx])
        n = 1
        for l, h in zip(lo, hi):
            n *= h - l
        covered += n
    return covered


def load(model: torch.nn.Module, optimizer: Optional[Any], path: str, strict: bool = True) -> Dict[str, Any]:
    """Restore ``model`` (+ fused optimizer state) in place from ``path``. Returns the manifest's ``extra`` dict."""
    with open(os.path.join(path, _MANIFEST)) as f:
        manifest = json.load(f)
    index = manifest["sharded"]
    blobs: Dict[int, Dict[str, torch.Tensor]] = {}
    dense = torch.load(os.path.join(path, "dense.pt"), map_location="cpu")
    targets: Dict[str, Any] = dict(model.state_dict())
    if optimizer is not None:
        targets.update({f"__optim__/{k}": v for k, v in _flatten_optim(optimizer.state_dict()).items()})
    missing: List[str] = []
    with torch.no_grad():
        for key, v in targets.items():
            if _is_sharded(v):
                if key not in index:
                    # saved unsharded (e.g. the table was data-parallel then): treat the dense tensor as one rectangle
                    if key in dense:
                        full = dense[key]
                        for t, off, size in _shards_of(v):
                            idx = tuple(slice(o, o + s) for o, s in zip(off, size))
                            t.copy_(full[idx])
                    else:
                        missing.append(key)
                    continue
                for t, off, size in _shards_of(v):
                    tmp = torch.empty(size, dtype=t.dtype)
                    n = _fill(tmp, off, size, index[key], blobs, path, _global_size(v))
                    if n != tmp.numel():
                        raise RuntimeError(f"checkpoint does not cover shard {off}+{size} of {key} ({n}/{tmp.numel()} elements)")
                    t.copy_(tmp)
            elif isinstance(v, torch.Tensor):
                if key in dense:
                    v.copy_(dense[key])
                elif key in index:  # saved sharded, now replicated: assemble the full tensor
                    tmp = torch.empty(list(v.shape), dtype=v.dtype)
                    _fill(tmp, [0] * v.dim(), list(v.shape), index[key], blobs, path)
                    v.copy_(tmp)
                else:
                    missing.append(key)
    if strict and missing:
        raise KeyError(f"checkpoint at {path} lacks {missing[:5]}{'...' if len(missing) > 5 else ''}")
    rank, world = _rank_world()
    if world > 1:
        dist.barrier()
    return manifest.get("extra", {})
