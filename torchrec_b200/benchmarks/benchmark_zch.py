"""Managed-collision (ZCH) remap throughput: sorted-ZCH ``MCHManagedCollisionModule`` profile + remap + eviction under a zipfian id
stream. Parity: reference ``distributed/benchmark/benchmark_zch`` (remap QPS, hit rate, eviction cost).

    python -m torchrec_b200.benchmarks.benchmark_zch --zch_size 100000 --batch_size 8192"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import torch

from ..modules.mc_modules import DistanceLFU_EvictionPolicy, LFU_EvictionPolicy, LRU_EvictionPolicy, MCHManagedCollisionModule
from ..sparse import JaggedTensor
from .base import BenchmarkResult, benchmark_func, cmd_conf


@dataclass
class ZchBenchConfig:
    zch_size: int = 10000
    id_space: int = 1000000
    batch_size: int = 2048
    eviction_interval: int = 4
    policy: str = "lfu"
    zipf_alpha: float = 1.1
    num_benchmarks: int = 12
    num_warmup: int = 4
    device: str = ""


def run(cfg: ZchBenchConfig) -> Dict[str, object]:
    device = torch.device(cfg.device or ("cuda" if torch.cuda.is_available() else "cpu"))
    pol = {"lfu": LFU_EvictionPolicy, "lru": LRU_EvictionPolicy, "distance_lfu": DistanceLFU_EvictionPolicy}[cfg.policy]()
    mc = MCHManagedCollisionModule(zch_size=cfg.zch_size, device=device, eviction_policy=pol, eviction_interval=cfg.eviction_interval, input_hash_size=cfg.id_space)
    mc.train()
    g = torch.Generator().manual_seed(0)
    ranks = torch.arange(1, cfg.id_space + 1, dtype=torch.float64)
    probs = ranks.pow(-cfg.zipf_alpha)
    probs /= probs.sum()
    stream: List[torch.Tensor] = [torch.multinomial(probs, cfg.batch_size, replacement=True, generator=g).to(device) for _ in range(8)]
    it = {"i": 0}
    hits = {"n": 0, "tot": 0}

    def step() -> None:
        ids = stream[it["i"] % len(stream)]
        it["i"] += 1
        feats = {"f": JaggedTensor(values=ids, lengths=torch.ones_like(ids, dtype=torch.int32))}
        mc.profile(feats)
        out = mc.remap(feats)["f"].values()
        hits["n"] += int((out < cfg.zch_size).sum())
        hits["tot"] += ids.numel()

    res = benchmark_func(f"zch_{cfg.policy}", step, cfg.num_benchmarks, cfg.num_warmup, device)
    qps = cfg.batch_size / (res.runtime_percentile(50) / 1e3)
    return {"result": res, "ids_per_s": qps, "in_zch_range": hits["n"] / max(hits["tot"], 1)}


@cmd_conf
def main(cfg: ZchBenchConfig) -> Dict[str, object]:
    r = run(cfg)
    print(r["result"])
    print(f"remap throughput {r['ids_per_s'] / 1e6:.2f} M ids/s, {100 * r['in_zch_range']:.1f}% of ids mapped into the collision-free range")
    return r


if __name__ == "__main__":
    main()
