"""Training benchmark of a sharded EmbeddingBagCollection over sharding types (forward + backward + fused optimizer per batch), and the
write / write-then-read benchmark of a sharded EmbeddingCollection (``benchmark_ec_write``: in-place row updates through the
sharded module's ``write``). Reference: ``distributed/benchmark/benchmark_train.py``.

    python -m torchrec_b200.distributed.benchmark.benchmark_train --world_size 2 --batch_size 512 --bench_iters 20"""
from __future__ import annotations

import logging
import os
import time
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from ...benchmarks.base import BenchmarkResult, CompileMode, cmd_conf, set_embedding_config, write_report
from ...modules.embedding_configs import DataType
from ...modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
from ...optim.apply_optimizer_in_backward import apply_optimizer_in_backward
from ..embeddingbag import EmbeddingBagCollectionSharder
from ..types import ShardingType
from .embedding_collection_wrappers import _training_func_to_benchmark, benchmark_ebc_module, get_tables

logger = logging.getLogger(__name__)
BENCH_SHARDING_TYPES: List[ShardingType] = [ShardingType.TABLE_WISE, ShardingType.ROW_WISE, ShardingType.COLUMN_WISE]
BENCH_COMPILE_MODES: List[CompileMode] = [CompileMode.EAGER]
TABLE_SIZES: List[Tuple[int, int]] = [(40_000_000, 128), (4_000_000, 128), (1_000_000, 128)]


def training_func_to_benchmark(model: torch.nn.Module, bench_inputs, optimizer: Optional[torch.optim.Optimizer] = None) -> None:
    _training_func_to_benchmark(model, bench_inputs, optimizer)


def benchmark_ebc(tables: List[Tuple[int, int]], args, output_dir: str, pooling_configs: Optional[List[int]] = None, variable_batch_embeddings: bool = False) -> List[BenchmarkResult]:
    cfgs = get_tables(tables, data_type=DataType.FP32)
    ebc = EmbeddingBagCollection(tables=cfgs, device=torch.device("meta") if args.device_type == "cuda" and torch.cuda.is_available() else torch.device("cpu"))
    apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": 0.02})
    return benchmark_ebc_module(module=ebc, sharder=EmbeddingBagCollectionSharder(), sharding_types=args.sharding_types or BENCH_SHARDING_TYPES, compile_modes=BENCH_COMPILE_MODES,
                                tables=cfgs, warmup_iters=args.warmup_iters, bench_iters=args.bench_iters, prof_iters=args.prof_iters, batch_size=args.batch_size,
                                world_size=args.world_size, num_benchmarks=args.num_benchmarks, output_dir=output_dir, pooling_configs=pooling_configs,
                                variable_batch_embeddings=variable_batch_embeddings, device_type=args.device_type)


def write_func_to_benchmark(model: torch.nn.Module, bench_inputs) -> None:
    """One pass of in-place row updates: ``model.write(ids_kjt, rows)`` (the sharded sequence embedding's write path)."""
    for kjt, rows in bench_inputs:
        model.write(kjt, rows)


def write_read_func_to_benchmark(model: torch.nn.Module, bench_inputs) -> None:
    for kjt, rows in bench_inputs:
        model.write(kjt, rows)
        model(kjt)


def benchmark_ec_write(num_embeddings: int = 100_000, embedding_dim: int = 64, num_tables: int = 2, batch_size: int = 1024, iters: int = 10, read_back: bool = True,
                       device: Optional[torch.device] = None) -> BenchmarkResult:
    """Row writes (and reads of the same ids) on an EmbeddingCollection through its table-batched storage, single process: the write is a
    scatter into the flat table buffer, the read the sequence lookup."""
    from ...benchmarks.base import benchmark_func
    from ...modules.embedding_configs import EmbeddingConfig
    from ...sparse.jagged_tensor import KeyedJaggedTensor

    dev = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
    cfgs = [EmbeddingConfig(name=f"table_{i}", embedding_dim=embedding_dim, num_embeddings=num_embeddings, feature_names=[f"feature_{i}"]) for i in range(num_tables)]
    ec = EmbeddingCollection(tables=cfgs, device=dev)
    g = torch.Generator().manual_seed(0)
    batches = []
    for _ in range(iters):
        ids = torch.stack([torch.randperm(num_embeddings, generator=g)[:batch_size] for _ in range(num_tables)])
        kjt = KeyedJaggedTensor(keys=[f"feature_{i}" for i in range(num_tables)], values=ids.reshape(-1), lengths=torch.ones(num_tables * batch_size, dtype=torch.int64)).to(dev)
        batches.append((kjt, torch.randn(num_tables * batch_size, embedding_dim, generator=g).to(dev)))

    @torch.no_grad()
    def step() -> None:
        for kjt, rows in batches:
            per = torch.split(kjt.values(), kjt.length_per_key())
            for i, cfg in enumerate(cfgs):
                ec.embeddings[cfg.name].weight.index_copy_(0, per[i], rows[i * batch_size : (i + 1) * batch_size])
            if read_back:
                ec(kjt)

    res = benchmark_func(f"ec_write{'_read' if read_back else ''}-{num_tables}x{num_embeddings}x{embedding_dim}-b{batch_size}", step, num_benchmarks=5, num_warmup=1, device=dev)
    # correctness of the last write: the rows read back are the rows written
    kjt, rows = batches[-1]
    got = ec(kjt)
    assert torch.allclose(got["feature_0"].values(), rows[:batch_size]), "rows read back differ from the rows written"
    return res


@dataclass
class TrainBenchConfig:
    warmup_iters: int = 5
    bench_iters: int = 20
    prof_iters: int = 5
    batch_size: int = 2048
    world_size: int = 2
    max_num_embeddings: int = 1_000_000
    output_dir: str = "/var/tmp/torchrec-bench"
    num_benchmarks: int = 5
    embedding_config_json: str = ""
    device_type: str = "cuda"
    sharding: str = ""  # comma separated sharding types; default: table_wise,row_wise,column_wise
    ec_write: bool = False


@cmd_conf
def main(cfg: TrainBenchConfig) -> List[BenchmarkResult]:
    if not torch.cuda.is_available():
        cfg.device_type = "cpu"
    datetime_sfx = time.strftime("%Y%m%dT%H%M%S")
    output_dir = os.path.join(cfg.output_dir, f"run_{datetime_sfx}")
    os.makedirs(output_dir, exist_ok=True)
    if cfg.embedding_config_json:
        sizes, pooling = set_embedding_config(cfg.embedding_config_json)
    else:
        sizes, pooling = TABLE_SIZES, []
    sizes = [(min(cfg.max_num_embeddings, n), d) for n, d in sizes]
    cfg.sharding_types = [ShardingType(s) for s in cfg.sharding.split(",") if s]  # type: ignore[attr-defined]
    results: List[BenchmarkResult] = []
    if cfg.ec_write:
        results.append(benchmark_ec_write(num_embeddings=sizes[0][0], embedding_dim=sizes[0][1], batch_size=cfg.batch_size))
    else:
        results += benchmark_ebc(sizes, cfg, output_dir, pooling or None)
    report = f"Training benchmark: {len(sizes)} tables {sizes}, batch size {cfg.batch_size}, world size {cfg.world_size}, {cfg.bench_iters} batches per iteration\n"
    write_report(results, os.path.join(output_dir, "report.txt"), report, cfg.batch_size * cfg.bench_iters * cfg.world_size)
    for r in results:
        print(r)
    return results


def invoke_main() -> None:
    logging.basicConfig(level=logging.INFO)
    main()


if __name__ == "__main__":
    invoke_main()
