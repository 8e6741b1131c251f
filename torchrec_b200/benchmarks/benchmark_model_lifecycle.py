"""Model lifecycle timings: build on meta device -> plan -> shard + materialise -> state_dict -> load_state_dict -> first step.
Parity: reference ``distributed/benchmark/benchmark_model_lifecycle`` (init / shard / checkpoint costs that dominate job start-up).

    torchrun --nproc-per-node 2 -m torchrec_b200.benchmarks.benchmark_model_lifecycle --num_tables 16"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import Dict

import torch
import torch.distributed as dist

from ..models.dlrm import DLRM, DLRMTrain
from ..modules.embedding_configs import EmbeddingBagConfig
from ..modules.embedding_modules import EmbeddingBagCollection
from ..optim.apply_optimizer_in_backward import apply_optimizer_in_backward
from ..optim.rowwise_adagrad import RowWiseAdagrad
from ..parallel.model_parallel import DistributedModelParallel
from ..parallel.planner import EmbeddingShardingPlanner, Topology
from ..parallel.types import ShardingEnv
from .base import cmd_conf


@dataclass
class LifecycleConfig:
    num_tables: int = 8
    num_embeddings: int = 100000
    embedding_dim: int = 64
    batch_size: int = 256
    device: str = ""


def run(cfg: LifecycleConfig) -> Dict[str, float]:
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        cuda = torch.cuda.is_available() and cfg.device != "cpu"
        dist.init_process_group("nccl" if cuda else "gloo", rank=int(os.environ.get("RANK", 0)), world_size=int(os.environ.get("WORLD_SIZE", 1)))
        created = True
    rank, world = dist.get_rank(), dist.get_world_size()
    cuda = torch.cuda.is_available() and cfg.device != "cpu"
    device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}") if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    t: Dict[str, float] = {}

    def lap(name: str, t0: float) -> float:
        if cuda:
            torch.cuda.synchronize()
        t[name] = (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    t0 = time.perf_counter()
    tables = [EmbeddingBagConfig(name=f"t{i}", embedding_dim=cfg.embedding_dim, num_embeddings=cfg.num_embeddings, feature_names=[f"f{i}"]) for i in range(cfg.num_tables)]
    ebc = EmbeddingBagCollection(tables, device=torch.device("meta"))
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.01})
    model = DLRMTrain(DLRM(ebc, dense_in_features=13, dense_arch_layer_sizes=[64, cfg.embedding_dim], over_arch_layer_sizes=[64, 1], dense_device=device))
    t0 = lap("build_meta_ms", t0)
    planner = EmbeddingShardingPlanner(topology=Topology(world_size=world, compute_device=device.type, local_world_size=world), batch_size=cfg.batch_size)
    plan = planner.collective_plan(model, pg=dist.group.WORLD)
    t0 = lap("plan_ms", t0)
    dmp = DistributedModelParallel(model, env=ShardingEnv.from_process_group(dist.group.WORLD), device=device, plan=plan)
    t0 = lap("shard_materialise_ms", t0)
    sd = dmp.state_dict()
    t0 = lap("state_dict_ms", t0)
    dmp.load_state_dict(sd)
    t0 = lap("load_state_dict_ms", t0)
    from ..datasets.random import RandomRecDataset

    ds = RandomRecDataset(keys=[f"f{i}" for i in range(cfg.num_tables)], batch_size=cfg.batch_size, hash_size=cfg.num_embeddings, ids_per_feature=2, num_dense=13,
                          manual_seed=rank, num_batches=2)
    batch = next(iter(ds)).to(device)
    loss, _ = dmp(batch)
    loss.backward()
    lap("first_step_ms", t0)
    if created:
        dist.destroy_process_group()
    return t


@cmd_conf
def main(cfg: LifecycleConfig) -> Dict[str, float]:
    t = run(cfg)
    if int(os.environ.get("RANK", 0)) == 0:
        for k, v in t.items():
            print(f"{k: <24} {v:10.1f} ms")
    return t


if __name__ == "__main__":
    main()
