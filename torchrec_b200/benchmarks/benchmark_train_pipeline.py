"""YAML-driven train-pipeline benchmark (reference torchrec/distributed/benchmark/benchmark_train_pipeline.py + yaml/*.yml).

    python -m torchrec_b200.benchmarks.benchmark_train_pipeline --yaml torchrec_b200/benchmarks/yaml/sparse_dist.yml
    torchrun --nproc-per-node 8 -m torchrec_b200.benchmarks.benchmark_train_pipeline --yaml ... 

Config keys (all optional): model {name: dlrm|deepfm, dense_arch, over_arch, embedding_dim}, tables {num, rows, pooling},
run {batch_size, steps, warmup, pipelines: [base, sparse_dist, sparse_dist_lite, fused_sparse_dist, semi_sync, prefetch],
sharding: table_wise|row_wise|column_wise|planner, dense_backend: torch|tcgen05}. Prints one JSON line per pipeline."""
from __future__ import annotations

import argparse
import json
import os
import time
from typing import Any, Dict, List

import torch
import torch.distributed as dist

DEFAULT: Dict[str, Any] = {
    "model": {"name": "dlrm", "dense_arch": [64, 32], "over_arch": [64, 1], "embedding_dim": 32},
    "tables": {"num": 8, "rows": 10000, "pooling": 4},
    "run": {"batch_size": 256, "steps": 10, "warmup": 3, "pipelines": ["base", "sparse_dist"], "sharding": "table_wise", "dense_backend": "torch"},
}


def _merge(a: Dict[str, Any], b: Dict[str, Any]) -> Dict[str, Any]:
    out = dict(a)
    for k, v in (b or {}).items():
        out[k] = _merge(a.get(k, {}), v) if isinstance(v, dict) else v
    return out


def build(cfg: Dict[str, Any], device: torch.device, world: int):
    from ..models.deepfm import SimpleDeepFMNN
    from ..models.dlrm import DLRM, DLRMTrain
    from ..modules.embedding_configs import EmbeddingBagConfig
    from ..modules.embedding_modules import EmbeddingBagCollection
    from ..ops import dense as _dense
    from ..optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from ..optim.keyed import CombinedOptimizer, KeyedOptimizerWrapper
    from ..optim.optimizers import in_backward_optimizer_filter
    from ..optim.rowwise_adagrad import RowWiseAdagrad
    from ..parallel import sharding_plan as sp
    from ..parallel.embeddingbag import EmbeddingBagCollectionSharder
    from ..parallel.model_parallel import DistributedModelParallel
    from ..parallel.types import ShardingPlan

    m, t, r = cfg["model"], cfg["tables"], cfg["run"]
    keys = [f"f{i}" for i in range(t["num"])]
    tables = [EmbeddingBagConfig(name=f"t{i}", embedding_dim=m["embedding_dim"], num_embeddings=t["rows"], feature_names=[keys[i]]) for i in range(t["num"])]
    ebc = EmbeddingBagCollection(tables, device=torch.device("meta"))
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.01})
    if m["name"] == "dlrm":
        model = DLRMTrain(DLRM(ebc, 13, list(m["dense_arch"]) + [m["embedding_dim"]], list(m["over_arch"]), dense_device=device))
        path = "model.sparse_arch.embedding_bag_collection"
    else:
        class _Train(torch.nn.Module):
            def __init__(self) -> None:
                super().__init__()
                self.model = SimpleDeepFMNN(13, ebc, 64, 16).to(device)

            def forward(self, batch):
                logits = self.model(batch.dense_features, batch.sparse_features).squeeze(-1)
                loss = torch.nn.functional.binary_cross_entropy(logits, batch.labels.float())
                return loss, (loss.detach(), logits.detach(), batch.labels)

        model, path = _Train(), "model.sparse_arch.embedding_bag_collection"
    if device.type == "cuda":
        _dense.set_dense_backend(r["dense_backend"])
    sharder = EmbeddingBagCollectionSharder()
    if r["sharding"] == "planner":
        plan = None
    else:
        gen = {"table_wise": lambda i: sp.table_wise(rank=i % world), "row_wise": lambda i: sp.row_wise(),
               "column_wise": lambda i: sp.column_wise(ranks=[(i + j) % world for j in range(min(world, 2))]) if world > 1 else sp.table_wise(rank=0)}[r["sharding"]]
        mp = sp.construct_module_sharding_plan(ebc, {tb.name: gen(i) for i, tb in enumerate(tables)}, sharder=sharder, world_size=world, local_size=world, device_type=device.type)
        plan = ShardingPlan({path: mp})
    dmp = DistributedModelParallel(model, device=device, plan=plan, sharders=[sharder])
    dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(dmp.named_parameters())), lambda p: torch.optim.SGD(p, lr=0.01))
    return dmp, CombinedOptimizer([dmp.fused_optimizer, dense_opt]), keys


def run(cfg: Dict[str, Any]) -> List[Dict[str, Any]]:
    from ..datasets.random import RandomRecDataset
    from ..parallel import train_pipeline as tp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cuda = torch.cuda.is_available()
    device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}") if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl" if cuda else "gloo")
    r, t = cfg["run"], cfg["tables"]
    classes = {"base": tp.TrainPipelineBase, "sparse_dist": tp.TrainPipelineSparseDist, "sparse_dist_lite": tp.TrainPipelineSparseDistLite,
               "fused_sparse_dist": tp.TrainPipelineFusedSparseDist, "semi_sync": tp.TrainPipelineSemiSync, "prefetch": tp.PrefetchTrainPipelineSparseDist}
    results = []
    for name in r["pipelines"]:
        dmp, opt, keys = build(cfg, device, world)
        ds = RandomRecDataset(keys, r["batch_size"], hash_sizes=[t["rows"]] * t["num"], ids_per_features=[t["pooling"]] * t["num"], num_dense=13, manual_seed=rank,
                              num_generated_batches=4, pin_memory=cuda)
        batches = ds.batch_generator._generated_batches
        pipe = classes[name](dmp, opt, device)
        total = r["warmup"] + r["steps"]
        it = iter(batches[i % len(batches)] for i in range(total + 4))
        for _ in range(r["warmup"]):
            pipe.progress(it)
        if cuda:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(r["steps"]):
            pipe.progress(it)
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res = {"pipeline": name, "ms_per_step": dt / r["steps"] * 1e3, "samples_per_s": r["batch_size"] * world * r["steps"] / dt, "world": world, "device": device.type,
               "sharding": r["sharding"], "batch_size": r["batch_size"]}
        if cuda:
            res["peak_mem_gb"] = torch.cuda.max_memory_allocated(device) / 1e9
        results.append(res)
        if rank == 0:
            print(json.dumps(res))
    return results


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--yaml", type=str, default=None)
    a = ap.parse_args()
    cfg = DEFAULT
    if a.yaml:
        import yaml

        with open(a.yaml) as f:
            cfg = _merge(DEFAULT, yaml.safe_load(f))
    run(cfg)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
