"""EmbeddingBagCollection variants: plain ``nn.EmbeddingBag`` EBC vs the table-batched fused EBC (fwd + bwd + fused optimizer) vs the
quantized inference EBC. Parity: reference ``benchmarks/ebc_benchmarks.py`` + ``distributed/benchmark/benchmark_train.py``.

    python -m torchrec_b200.benchmarks.benchmark_ebc --num_tables 8 --num_embeddings 100000 --batch_size 4096"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch

from ..modules.embedding_configs import DataType, EmbeddingBagConfig
from ..modules.embedding_modules import EmbeddingBagCollection
from ..modules.fused_embedding_modules import FusedEmbeddingBagCollection
from ..quant.embedding_modules import EmbeddingBagCollection as QuantEmbeddingBagCollection
from ..sparse import KeyedJaggedTensor
from .base import BenchmarkResult, benchmark_func, cmd_conf


@dataclass
class EbcBenchConfig:
    num_tables: int = 4
    num_embeddings: int = 10000
    embedding_dim: int = 64
    batch_size: int = 512
    pooling_factor: int = 8
    num_benchmarks: int = 10
    num_warmup: int = 3
    device: str = ""
    profile_dir: str = ""
    quant_dtype: str = "int8"


def _batch(cfg: EbcBenchConfig, device: torch.device, seed: int = 0) -> KeyedJaggedTensor:
    g = torch.Generator().manual_seed(seed)
    F, B, L = cfg.num_tables, cfg.batch_size, cfg.pooling_factor
    lengths = torch.full((F * B,), L, dtype=torch.int32)
    values = torch.randint(0, cfg.num_embeddings, (F * B * L,), generator=g)
    return KeyedJaggedTensor(keys=[f"f{i}" for i in range(F)], values=values, lengths=lengths).to(device)


def run(cfg: EbcBenchConfig) -> List[BenchmarkResult]:
    device = torch.device(cfg.device or ("cuda" if torch.cuda.is_available() else "cpu"))
    tables = [EmbeddingBagConfig(name=f"t{i}", embedding_dim=cfg.embedding_dim, num_embeddings=cfg.num_embeddings, feature_names=[f"f{i}"]) for i in range(cfg.num_tables)]
    kjt = _batch(cfg, device)
    out: List[BenchmarkResult] = []

    ebc = EmbeddingBagCollection(tables, device=device)
    opt = torch.optim.SGD(ebc.parameters(), lr=0.1)

    def step_plain() -> None:
        opt.zero_grad()
        ebc(kjt).values().sum().backward()
        opt.step()

    out.append(benchmark_func("ebc_fwd_bwd_sgd", step_plain, cfg.num_benchmarks, cfg.num_warmup, device, cfg.profile_dir))

    fused = FusedEmbeddingBagCollection(tables, torch.optim.SGD, {"lr": 0.1}, device=device)

    def step_fused() -> None:
        fused(kjt).values().sum().backward()

    out.append(benchmark_func("fused_ebc_fwd_bwd_sgd", step_fused, cfg.num_benchmarks, cfg.num_warmup, device, cfg.profile_dir))

    float_ebc = EmbeddingBagCollection(tables, device=torch.device("cpu"))
    dt = {"int8": DataType.INT8, "int4": DataType.INT4, "fp16": DataType.FP16}[cfg.quant_dtype]
    float_ebc.qconfig = torch.ao.quantization.QConfig(activation=torch.ao.quantization.PlaceholderObserver.with_args(dtype=torch.float),
                                                      weight=torch.ao.quantization.PlaceholderObserver.with_args(dtype={DataType.INT8: torch.qint8, DataType.INT4: torch.quint4x2,
                                                                                                                 DataType.FP16: torch.float16}[dt]))
    q = QuantEmbeddingBagCollection.from_float(float_ebc)
    if device.type == "cuda":
        q = q.to(device)
    with torch.no_grad():
        out.append(benchmark_func(f"quant_ebc_{cfg.quant_dtype}_fwd", lambda: q(kjt), cfg.num_benchmarks, cfg.num_warmup, device, cfg.profile_dir))
    return out


@cmd_conf
def main(cfg: EbcBenchConfig) -> List[BenchmarkResult]:
    res = run(cfg)
    for r in res:
        print(r)
    return res


if __name__ == "__main__":
    main()
