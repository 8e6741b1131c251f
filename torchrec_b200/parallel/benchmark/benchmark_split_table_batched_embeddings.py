"""Operator benchmark of the table-batched embedding kernel (forward, and forward + backward with the fused optimizer) over embedding
dimensions. Reference: ``distributed/benchmark/benchmark_split_table_batched_embeddings.py`` (FBGEMM's SplitTableBatchedEmbeddingBagsCodegen);
here the operator is this framework's ``TableBatchedEmbeddingBags`` (``ops/csrc/tbe_fwd.cu`` / ``tbe_bwd.cu``).

    python -m torchrec_b200.distributed.benchmark.benchmark_split_table_batched_embeddings --num_embeddings 1000000 --embedding_dim 128"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List

import torch

from ...benchmarks.base import BenchmarkResult, benchmark_inputs_func, cmd_conf
from ...ops.tbe import OptimType, PoolingMode, TableBatchedEmbeddingBags
from ...sparse.jagged_tensor import KeyedJaggedTensor


@dataclass
class TbeBenchConfig:
    num_embeddings: int = 100000
    embedding_dim: int = 128  # 0: sweep 4 .. 1024
    num_tables: int = 4
    batch_size: int = 4096
    bag_size: int = 10
    num_benchmarks: int = 10
    backward: bool = True
    device: str = ""
    profile_dir: str = ""


def op_bench(num_embeddings: int, embedding_dim: int, num_tables: int, batch_size: int, bag_size: int, num_benchmarks: int = 10, backward: bool = True, device: str = "",
             profile_dir: str = "") -> BenchmarkResult:
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    emb = TableBatchedEmbeddingBags([(num_embeddings, embedding_dim)] * num_tables, optimizer=OptimType.EXACT_ADAGRAD, learning_rate=0.1, eps=0.1,
                                    pooling_mode=PoolingMode.SUM, device=dev)
    g = torch.Generator().manual_seed(0)
    lengths = torch.full((num_tables * batch_size,), bag_size, dtype=torch.int64)
    kjt = KeyedJaggedTensor(keys=[f"feature_{i}" for i in range(num_tables)], values=torch.randint(0, num_embeddings, (int(lengths.sum()),), generator=g), lengths=lengths).to(dev)

    def _func_to_benchmark(kjts: List[Dict[str, KeyedJaggedTensor]], model: torch.nn.Module) -> torch.Tensor:
        k = kjts[0]["feature"]
        out = model(k.values(), k.offsets())
        if backward:
            out.sum().backward()
        return out

    name = f"TableBatchedEmbeddingBags-{num_embeddings}-{embedding_dim}-{num_tables}-{batch_size}-{bag_size}" + ("-fwdbwd" if backward else "-fwd")
    res = benchmark_inputs_func(name=name, rank=0, world_size=1, func_to_benchmark=_func_to_benchmark, bench_inputs=[{"feature": kjt}], prof_inputs=[{"feature": kjt}],
                                benchmark_func_kwargs={"model": emb}, num_profiles=3, num_benchmarks=num_benchmarks, profile_dir=profile_dir, device_type=dev.type,
                                sample_count=batch_size)
    print(res)
    return res


@cmd_conf
def main(cfg: TbeBenchConfig) -> List[BenchmarkResult]:
    dims = [4, 8, 16, 32, 64, 128, 256, 512, 1024] if cfg.embedding_dim == 0 else [cfg.embedding_dim]
    return [op_bench(cfg.num_embeddings, d, cfg.num_tables, cfg.batch_size, cfg.bag_size, cfg.num_benchmarks, cfg.backward, cfg.device, cfg.profile_dir) for d in dims]


if __name__ == "__main__":
    main()
