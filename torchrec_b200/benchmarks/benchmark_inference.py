"""Serving path latency: quantize a trained DLRM (per-table INT8 / INT4), shard it over the local devices with the inference planner,
time batched predict calls. Parity: reference ``distributed/benchmark/benchmark_inference`` + ``inference/dlrm_predict.py``.

    python -m torchrec_b200.benchmarks.benchmark_inference --batch_size 512 --world_size 1"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch

from ..inference.modules import quantize_inference_model, shard_quant_model
from ..models.dlrm import DLRM
from ..modules.embedding_configs import DataType, EmbeddingBagConfig
from ..modules.embedding_modules import EmbeddingBagCollection
from ..sparse import KeyedJaggedTensor
from .base import BenchmarkResult, benchmark_func, cmd_conf


@dataclass
class InferenceBenchConfig:
    num_tables: int = 8
    num_embeddings: int = 100000
    embedding_dim: int = 64
    batch_size: int = 256
    pooling_factor: int = 4
    world_size: int = 1
    weight_dtype: str = "int8"
    num_benchmarks: int = 20
    num_warmup: int = 5
    profile_dir: str = ""


def run(cfg: InferenceBenchConfig) -> List[BenchmarkResult]:
    cuda = torch.cuda.is_available()
    device = torch.device("cuda:0" if cuda else "cpu")
    tables = [EmbeddingBagConfig(name=f"t{i}", embedding_dim=cfg.embedding_dim, num_embeddings=cfg.num_embeddings, feature_names=[f"f{i}"]) for i in range(cfg.num_tables)]
    model = DLRM(EmbeddingBagCollection(tables, device=torch.device("cpu")), dense_in_features=13, dense_arch_layer_sizes=[64, cfg.embedding_dim],
                 over_arch_layer_sizes=[64, 1], dense_device=torch.device("cpu"))
    model.eval()
    dt = {"int8": DataType.INT8, "int4": DataType.INT4, "fp16": DataType.FP16}[cfg.weight_dtype]
    qmodel = quantize_inference_model(model, quantization_dtype=dt)
    sharded, _plan = shard_quant_model(qmodel, world_size=cfg.world_size, compute_device=device.type, sharding_device="cpu")
    sharded = sharded.to(device) if cuda else sharded
    g = torch.Generator().manual_seed(0)
    F, B, L = cfg.num_tables, cfg.batch_size, cfg.pooling_factor
    kjt = KeyedJaggedTensor(keys=[f"f{i}" for i in range(F)], values=torch.randint(0, cfg.num_embeddings, (F * B * L,), generator=g),
                            lengths=torch.full((F * B,), L, dtype=torch.int32)).to(device)
    dense = torch.randn(B, 13, generator=g).to(device)
    with torch.no_grad():
        res = benchmark_func(f"dlrm_predict_{cfg.weight_dtype}_w{cfg.world_size}", lambda: sharded(dense, kjt), cfg.num_benchmarks, cfg.num_warmup, device, cfg.profile_dir)
    return [res]


@cmd_conf
def main(cfg: InferenceBenchConfig) -> List[BenchmarkResult]:
    res = run(cfg)
    for r in res:
        print(r, f"| {cfg.batch_size / (r.runtime_percentile(50) / 1e3):.0f} samples/s")
    return res


if __name__ == "__main__":
    main()
