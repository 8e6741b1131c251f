"""Benchmark plumbing for embedding collections: wrappers with plain outputs, table / input generators and
``benchmark_ebc_module`` - one sharded run per (sharding type, compile mode), every rank a process (NCCL on GPUs, gloo on CPU), timed
with the harness of ``benchmarks/base.py`` (reference ``distributed/benchmark/embedding_collection_wrappers.py``)."""
from __future__ import annotations

import logging
import multiprocessing
import os
import tempfile
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from ...benchmarks.base import BenchmarkResult, CompileMode, benchmark_model_with_warmup
from ...modules.embedding_configs import DataType, EmbeddingBagConfig, EmbeddingConfig
from ...sparse.jagged_tensor import JaggedTensor, KeyedJaggedTensor, KeyedTensor
from ..planner.types import ParameterConstraints
from ..types import ShardingType

logger = logging.getLogger(__name__)


class ECWrapper(nn.Module):
    """``forward(kjt) -> Dict[str, JaggedTensor]`` of a sequence embedding module (sharded modules return an awaitable: waited here)."""

    def __init__(self, module: nn.Module) -> None:
        super().__init__()
        self._module = module

    def forward(self, input: KeyedJaggedTensor) -> Dict[str, JaggedTensor]:
        out = self._module(input)
        return out.wait() if hasattr(out, "wait") else out


class EBCWrapper(nn.Module):
    """``forward(kjt) -> KeyedTensor`` of a pooled embedding module."""

    def __init__(self, module: nn.Module) -> None:
        super().__init__()
        self._module = module

    def forward(self, input: KeyedJaggedTensor) -> KeyedTensor:
        out = self._module(input)
        return out.wait() if hasattr(out, "wait") else out


def _default_func_to_benchmark(model: nn.Module, bench_inputs: List[KeyedJaggedTensor]) -> None:
    with torch.inference_mode():
        for x in bench_inputs:
            model(x)


def _training_func_to_benchmark(model: nn.Module, bench_inputs: List[KeyedJaggedTensor], optimizer: Optional[torch.optim.Optimizer] = None) -> None:
    for x in bench_inputs:
        out = model(x)
        vals = out.values() if isinstance(out, KeyedTensor) else torch.cat([jt.values().reshape(-1) for jt in out.values()])
        vals.sum().backward()
        if optimizer is not None:
            optimizer.step()
            optimizer.zero_grad()


def get_tables(table_sizes: List[Tuple[int, int]], is_pooled: bool = True, data_type: DataType = DataType.INT8) -> Union[List[EmbeddingBagConfig], List[EmbeddingConfig]]:
    cls = EmbeddingBagConfig if is_pooled else EmbeddingConfig
    return [cls(num_embeddings=n, embedding_dim=d, name=f"table_{i}", feature_names=[f"feature_{i}"], data_type=data_type) for i, (n, d) in enumerate(table_sizes)]


def _get_inputs(tables, batch_size: int, n: int, pooling_configs: Optional[List[int]], device: torch.device, seed: int, variable_batch: bool = False) -> List[KeyedJaggedTensor]:
    g = torch.Generator().manual_seed(seed)
    keys = [f for t in tables for f in t.feature_names]
    out = []
    for _ in range(n):
        lens, vals, strides = [], [], []
        for i, t in enumerate(tables):
            pool = pooling_configs[i] if pooling_configs else 10
            b = int(torch.randint(max(batch_size // 2, 1), batch_size + 1, (1,), generator=g)) if variable_batch else batch_size
            ln = torch.randint(0, 2 * pool + 1, (b,), generator=g)
            lens.append(ln)
            vals.append(torch.randint(0, t.num_embeddings, (int(ln.sum()),), generator=g))
            strides.append([b])
        out.append(KeyedJaggedTensor(keys=keys, values=torch.cat(vals), lengths=torch.cat(lens), stride_per_key_per_rank=strides if variable_batch else None).to(device))
    return out


def _benchmark_type_name(compile_mode: CompileMode, sharding_type: ShardingType) -> str:
    return f"{sharding_type.value}-{compile_mode.value}" if sharding_type is not None else f"unsharded-{compile_mode.value}"


def _init_module_and_run_benchmark(ctx, module_bytes: bytes, sharder, sharding_type: Optional[ShardingType], compile_mode: CompileMode, tables, warmup_iters: int,
                                   bench_iters: int, prof_iters: int, batch_size: int, num_benchmarks: int, output_dir: str, func_to_benchmark, benchmark_func_kwargs,
                                   pooling_configs, variable_batch_embeddings: bool, constraints, local_world_size: Optional[int], result_path: str) -> None:
    """One rank of one benchmark: shard the module for ``sharding_type`` (every table), time it, rank 0 stores the result."""
    import io
    import pickle

    from .. import sharding_plan as sp
    from ..model_parallel import DistributedModelParallel
    from ..types import ShardingPlan

    module = torch.load(io.BytesIO(module_bytes), weights_only=False)
    dev = ctx.device
    is_pooled = hasattr(module, "embedding_bag_configs")
    training = not type(module).__module__.startswith("torchrec_b200.quant")
    wrapped = (EBCWrapper if is_pooled else ECWrapper)(module)
    if sharding_type is not None:
        gens = {ShardingType.TABLE_WISE: lambda i: sp.table_wise(rank=i % ctx.world_size), ShardingType.ROW_WISE: lambda i: sp.row_wise(),
                ShardingType.COLUMN_WISE: lambda i: sp.column_wise(ranks=list(range(min(ctx.world_size, 2)))), ShardingType.DATA_PARALLEL: lambda i: sp.data_parallel(),
                ShardingType.TABLE_ROW_WISE: lambda i: sp.table_row_wise(host_index=0)}
        per_table = {t.name: gens[sharding_type](i) for i, t in enumerate(tables)}
        plan = sp.construct_module_sharding_plan(module, per_table, sharder=sharder, world_size=ctx.world_size, local_size=local_world_size or ctx.world_size, device_type=dev.type)
        wrapped = DistributedModelParallel(wrapped, device=dev, plan=ShardingPlan({"_module": plan}), sharders=[sharder])
    else:
        wrapped = wrapped.to(dev)
    wrapped.train(training)
    inputs = lambda n, seed: _get_inputs(tables, batch_size, n, pooling_configs, dev, 1000 * ctx.rank + seed, variable_batch_embeddings)  # noqa: E731
    if func_to_benchmark is None:
        func_to_benchmark = _training_func_to_benchmark if training else _default_func_to_benchmark
    res = benchmark_model_with_warmup(name=_benchmark_type_name(compile_mode, sharding_type), model=wrapped, warmup_inputs=inputs(warmup_iters, 1), bench_inputs=inputs(bench_iters, 2),
                                      prof_inputs=inputs(prof_iters, 3), world_size=ctx.world_size, output_dir=output_dir, num_benchmarks=num_benchmarks,
                                      func_to_benchmark=func_to_benchmark, benchmark_func_kwargs=benchmark_func_kwargs, rank=ctx.rank, enable_logging=False, device_type=dev.type)
    if ctx.rank == 0:
        with open(result_path, "wb") as f:
            pickle.dump(res, f)


def benchmark_ebc_module(module: nn.Module, sharder, sharding_types: List[ShardingType], compile_modes: List[CompileMode], tables, warmup_iters: int = 20, bench_iters: int = 500,
                         prof_iters: int = 20, batch_size: int = 2048, world_size: int = 2, num_benchmarks: int = 5, output_dir: str = "", benchmark_unsharded: bool = False,
                         func_to_benchmark: Optional[Callable[..., None]] = None, benchmark_func_kwargs: Optional[Dict[str, Any]] = None, pooling_configs: Optional[List[int]] = None,
                         variable_batch_embeddings: bool = False, device_type: str = "cuda", pod_size: Optional[int] = None, local_world_size: Optional[int] = None,
                         constraints: Optional[Dict[str, ParameterConstraints]] = None) -> List[BenchmarkResult]:
    """Benchmark an (unsharded) embedding collection under every requested sharding type: ``world_size`` processes per run; one timed
    iteration = a pass over ``bench_iters`` random batches (forward + backward + fused optimizer for float modules, inference forward for
    quantized ones, or ``func_to_benchmark(model, inputs, **kwargs)``). ``benchmark_unsharded`` adds a single-process run."""
    import io
    import pickle

    from ...utils.multiprocess import run_multi_process

    logger.info("Warmup iterations: %d, Benchmark iterations: %d, Profile iterations: %d, Batch Size: %d, World Size: %d, Number of Benchmarks: %d, Output Directory: %s",
                warmup_iters, bench_iters, prof_iters, batch_size, world_size, num_benchmarks, output_dir)
    cuda = device_type == "cuda" and torch.cuda.is_available() and torch.cuda.device_count() >= world_size
    buf = io.BytesIO()
    torch.save(module, buf)
    results: List[BenchmarkResult] = []
    runs: List[Tuple[Optional[ShardingType], int]] = [(st, world_size) for st in sharding_types] + ([(None, 1)] if benchmark_unsharded else [])
    for compile_mode in compile_modes:
        for st, W in runs:
            with tempfile.TemporaryDirectory() as tmp:
                path = os.path.join(tmp, "result.pkl")
                run_multi_process(_init_module_and_run_benchmark, world_size=W, backend="nccl" if cuda else "gloo", timeout=1800.0, module_bytes=buf.getvalue(), sharder=sharder,
                                  sharding_type=st, compile_mode=compile_mode, tables=tables, warmup_iters=warmup_iters, bench_iters=bench_iters, prof_iters=prof_iters,
                                  batch_size=batch_size, num_benchmarks=num_benchmarks, output_dir=output_dir, func_to_benchmark=func_to_benchmark,
                                  benchmark_func_kwargs=benchmark_func_kwargs, pooling_configs=pooling_configs, variable_batch_embeddings=variable_batch_embeddings,
                                  constraints=constraints, local_world_size=local_world_size, result_path=path)
                with open(path, "rb") as f:
                    results.append(pickle.load(f))
    return results
