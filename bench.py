#!/usr/bin/env python
"""Headline benchmark: DLRM (Criteo-1TB shape, 26 sparse + 13 dense, MLPerf arch) training
throughput in samples/s, whole job, device timed, max over ranks.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference ...        # unmodified reference from baseline/_ref (if installable)

Prints ONE JSON line on rank 0 (see README / DESIGN.md for the field contract).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from typing import Any, Dict, List, Optional

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Criteo-1TB categorical cardinalities with the MLPerf 40M row cap (26 features)
CRITEO_1TB_40M = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10, 2208, 11938, 155, 4, 976, 14,
                  39979771, 25641295, 39664984, 585935, 12972, 108, 36]
BASELINE_SAMPLES_PER_SEC = {1: 50_000.0, 8: 350_000.0}  # reference published numbers (A100), BASELINE.md


def parse_args() -> argparse.Namespace:
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    p.add_argument("--batch-size", type=int, default=int(os.environ.get("TRB_BENCH_BATCH", 32768)), help="per-GPU batch (weak scaling)")
    p.add_argument("--embedding-dim", type=int, default=128)
    p.add_argument("--dense-arch", type=str, default="512,256,128")
    p.add_argument("--over-arch", type=str, default="1024,1024,512,256,1")
    p.add_argument("--row-cap", type=int, default=int(os.environ.get("TRB_BENCH_ROW_CAP", 40_000_000)))
    p.add_argument("--pooling", type=int, default=1, help="ids per sparse feature (Criteo is one-hot)")
    p.add_argument("--lr", type=float, default=0.01)
    p.add_argument("--sharding", type=str, default="auto", choices=["auto", "table_wise", "row_wise", "column_wise", "planner"],
                   help="auto: greedy table-wise on 1 GPU, table-wise + data-parallel tiny tables (what EmbeddingShardingPlanner also picks) on N > 1")
    p.add_argument("--dense-backend", type=str, default=os.environ.get("TRB_DENSE_BACKEND", "auto"))
    p.add_argument("--transport", type=str, default=os.environ.get("TRB_TRANSPORT", "auto"), help="auto | p2p | nccl")
    p.add_argument("--dp-rows", type=int, default=int(os.environ.get("TRB_BENCH_DP_ROWS", 2000)), help="tables with at most this many rows are data-parallel when N > 1 (0 = all table-wise)")
    p.add_argument("--cuda-graphs", type=int, default=int(os.environ.get("TRB_BENCH_GRAPHS", -1)),
                   help="1 (= -1, the default): replay the dense sub-modules as CUDA graphs (host enqueue 1.6-2.6 ms -> 0.74 ms per step; "
                        "end to end 15.4 M -> 17.7 M samples/s on one GPU); 0: eager")
    p.add_argument("--profile-host", action="store_true", help="cProfile 10 extra steps on rank 0 (stderr)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--trace-e2e", type=str, default="", help="diagnostics: torch.profiler trace (chrome json + op table) of 6 extra pipeline steps")
    p.add_argument("--num-host-batches", type=int, default=8)
    return p.parse_args()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int) -> None:
        self.gpu = gpu_index
        self.proc: Optional[subprocess.Popen] = None
        self.lines: List[str] = []
        self._t: Optional[threading.Thread] = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def reader() -> None:
            assert self.proc is not None and self.proc.stdout is not None
            for line in self.proc.stdout:
                self.lines.append((time.time(), line.strip()))

        self._t = threading.Thread(target=reader, daemon=True)
        self._t.start()

    def mark_start(self) -> None:
        """The sampler is started before warm-up (nvidia-smi takes longer to start than a short timed region lasts);
        only samples that arrive between mark_start() and stop() are reported."""
        self.t0 = time.time()

    def stop(self) -> Dict[str, Any]:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        inside = [ln for ts, ln in self.lines if t0 <= ts <= t1 + 0.11]
        window = "timed"
        if not inside:  # region shorter than one sampling period: fall back to the closest samples taken under (warm-up) load
            inside = [ln for ts, ln in self.lines][-3:]
            window = "warmup+timed"
        self.window = window
        for ln in inside:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "window": getattr(self, "window", "timed")}


def reference_arm(args: argparse.Namespace) -> None:
    """Run the UNMODIFIED reference from baseline/_ref if it is importable; otherwise say why."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    why = None
    if not os.path.isdir(ref):
        why = "baseline/_ref missing (reference not installed: fbgemm_gpu wheel unavailable offline, see DESIGN.md)"
    else:
        sys.path.insert(0, ref)
        try:
            import torchrec  # noqa: F401
        except Exception as e:  # fbgemm_gpu hard import
            why = f"reference import failed: {type(e).__name__}: {str(e).splitlines()[0][:160]}"
    if why is not None:
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": why}))
        return
    from baseline.run_reference import run as run_reference  # type: ignore

    run_reference(args)


def build_ours(args: argparse.Namespace, device, rank: int, world: int):
    import torch
    import torch.distributed as dist

    from torchrec_b200.models.dlrm import DLRM, DLRMTrain
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.ops import dense as _dense
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.keyed import CombinedOptimizer, KeyedOptimizerWrapper
    from torchrec_b200.optim.optimizers import in_backward_optimizer_filter
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan

    hashes = [min(h, args.row_cap) for h in CRITEO_1TB_40M]
    keys = [f"cat_{i}" for i in range(26)]
    D = args.embedding_dim
    tables = [EmbeddingBagConfig(name=f"t_{keys[i]}", embedding_dim=D, num_embeddings=hashes[i], feature_names=[keys[i]]) for i in range(26)]
    ebc = EmbeddingBagCollection(tables=tables, device=torch.device("meta"))
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": args.lr, "eps": 1e-8})
    dense_arch = [int(x) for x in args.dense_arch.split(",")]
    over_arch = [int(x) for x in args.over_arch.split(",")]
    model = DLRMTrain(DLRM(ebc, 13, dense_arch, over_arch, dense_device=device))

    backend = args.dense_backend
    if backend == "auto":
        backend = os.environ.get("TRB_DENSE_DEFAULT", "tcgen05")  # bf16 tensor-core dense path is the product
    _dense.set_dense_backend(backend)

    # bf16 pooled embeddings (half the NVLink / HBM bytes) when the dense arch computes in bf16
    fused_params = {"output_dtype": torch.bfloat16} if backend == "tcgen05" else None
    sharder = EmbeddingBagCollectionSharder(fused_params=fused_params)
    if args.sharding == "auto":
        args.sharding = "table_wise"
    if args.sharding == "planner":
        plan = None
    else:
        if args.sharding == "table_wise":
            # greedy balance: bytes first (big tables spread), then lookups per rank
            load = [[0.0, 0] for _ in range(world)]
            gens = {}
            order = sorted(range(26), key=lambda i: -hashes[i])
            tot = float(sum(hashes))
            if world > 1 and args.dp_rows > 0:
                # tiny tables are replicated (data parallel, dense gradient all-reduced) like the reference planner would place
                # them: no all-to-all traffic for their features and the model-parallel tables divide evenly over the ranks
                tiny = [i for i in order if hashes[i] <= args.dp_rows]
                keep = (26 - len(tiny)) % world  # keep the MP table count a multiple of the world size when possible
                tiny = tiny[keep:] if keep and len(tiny) > keep else tiny
                for i in tiny:
                    gens[tables[i].name] = sp.data_parallel()
                order = [i for i in order if i not in tiny]
            for i in order:
                r = min(range(world), key=lambda r: (load[r][0] / tot * world + load[r][1] / 26.0 * world, r))
                load[r][0] += hashes[i]
                load[r][1] += 1
                gens[tables[i].name] = sp.table_wise(rank=r)
        elif args.sharding == "row_wise":
            gens = {t.name: sp.row_wise() for t in tables}
        else:
            gens = {t.name: (sp.column_wise(ranks=[(i + j) % world for j in range(min(world, 4))]) if world > 1 else sp.table_wise(rank=0)) for i, t in enumerate(tables)}
        mplan = sp.construct_module_sharding_plan(ebc, gens, sharder=sharder, world_size=world, local_size=world, device_type="cuda")
        plan = ShardingPlan({"model.sparse_arch.embedding_bag_collection": mplan})
    # data-parallel wrapping is deferred (main() calls dmp.init_data_parallel()) so that CUDA graphs of the dense sub-modules can
    # be captured first: DDP keeps the parameters' AccumulateGrad nodes alive on the default stream, which a capture may not touch
    dmp = DistributedModelParallel(model, device=device, plan=plan, sharders=[sharder], init_data_parallel=False)
    dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(dmp.named_parameters())), lambda params: torch.optim.SGD(params, lr=args.lr))
    opt = CombinedOptimizer([dmp.fused_optimizer, dense_opt])
    return dmp, opt, keys, hashes, backend


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from torchrec_b200.datasets.random import RandomRecDataset
    from torchrec_b200.ops import _lib
    from torchrec_b200.parallel.train_pipeline import TrainPipelineSparseDist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    if os.environ.get("TRB_BENCH_BIND_NUMA", "1") != "0":
        from torchrec_b200.utils.affinity import bind_to_gpu_numa

        _bound = bind_to_gpu_numa(local_rank)  # pinned staging buffers + launch threads on the GPU's socket (H2D does not cross sockets)
        globals()["_CPU_BINDING"] = f"{len(_bound)} NUMA-local cpus" if _bound else "unchanged"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=device)
    _lib.lib()

    dmp, opt, keys, hashes, dense_backend = build_ours(args, device, rank, world)
    B = args.batch_size
    ds = RandomRecDataset(keys, B, hash_sizes=hashes, ids_per_features=[args.pooling] * 26, num_dense=13, manual_seed=1234 + rank,
                          num_generated_batches=args.num_host_batches, pin_memory=True)
    host_batches = ds.batch_generator._generated_batches
    dev_batches = [b.to(device) for b in host_batches]
    torch.cuda.synchronize()

    if args.cuda_graphs < 0:
        args.cuda_graphs = 1
    if not args.cuda_graphs:
        dmp.init_data_parallel()
    if args.cuda_graphs:
        # the dense sub-modules (bottom MLP, interaction + top MLP + head) replay as CUDA graphs: their ~50 launches and ~100 ATen
        # calls per step made the step launch-bound once DDP / NVLink dists were added (host enqueue 2.0 ms vs 2.1 ms of GPU time)
        inner = dmp.module.model
        with torch.no_grad():
            sample_emb = inner.sparse_arch(dev_batches[0].sparse_features)
        inner.capture_dense_graphs(dev_batches[0].dense_features, sample_emb)
        torch.cuda.synchronize()
        dmp.init_data_parallel()

    def step(batch) -> "torch.Tensor":
        opt.zero_grad()
        loss, _ = dmp(batch)
        loss.backward()
        opt.step()
        return loss

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (kernel-side number) ----------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        step(dev_batches[i % len(dev_batches)])
    barrier()
    sampler.mark_start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        loss = step(dev_batches[i % len(dev_batches)])
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps  # CPU time to ENQUEUE a step: >= ms_per_step means launch-bound
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    clocks = sampler.stop()
    if world > 1:  # worst GPU of the job: lowest median SM clock, union of throttle reasons
        allc: List[Any] = [None] * world
        dist.all_gather_object(allc, clocks)
        ok = [c for c in allc if c.get("sm_mhz") is not None]
        if ok:
            clocks = dict(min(ok, key=lambda c: c["sm_mhz"]))
            clocks["reasons"] = sorted({r for c in allc for r in c.get("reasons", [])})
            clocks["samples"] = sum(c.get("samples", 0) for c in ok)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = B * world * args.steps / (ms_max / 1e3)

    # ---------------- end-to-end through the public pipeline API ------------------------------------
    e2e: Optional[Dict[str, Any]] = None
    if not args.no_e2e:
        pipe = TrainPipelineSparseDist(dmp, opt, device)

        def host_iter(n: int):
            for i in range(n):
                yield host_batches[i % len(host_batches)]

        loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        total = max(args.warmup, 3) + args.steps
        it = host_iter(total + 2)
        for _ in range(max(args.warmup, 3)):
            out = pipe.progress(it)
        barrier()
        e0.record()
        t_e2e0 = time.perf_counter()
        for _ in range(args.steps):
            out = pipe.progress(it)
            loss_host.copy_(out[0].detach().reshape(1), non_blocking=True)  # D2H read of the step's loss
        e2e_host_ms = (time.perf_counter() - t_e2e0) * 1e3 / args.steps
        e1.record()
        barrier()
        ms2 = e0.elapsed_time(e1)
        t2 = torch.tensor([ms2], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * args.steps / (float(t2.item()) / 1e3), "unit": "samples/s",
               "h2d_bytes_per_step": host_batches[0].nbytes(), "d2h_bytes_per_step": 4, "ms_per_step": float(t2.item()) / args.steps, "host_enqueue_ms_per_step": e2e_host_ms,
               "loss": float(loss_host.item())}

    if args.trace_e2e and not args.no_e2e:
        from torch.profiler import ProfilerActivity, profile

        it3 = host_iter(6 + 4)
        for _ in range(2):
            pipe.progress(it3)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(4):
                out = pipe.progress(it3)
                loss_host.copy_(out[0].detach().reshape(1), non_blocking=True)
            torch.cuda.synchronize()
        if rank == 0:
            prof.export_chrome_trace(args.trace_e2e)
            with open(args.trace_e2e + ".txt", "w") as f:
                f.write(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=70))
                f.write("\n\n")
                f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
    if args.profile_host:
        # where does the host time of a step go? (diagnostics only; every rank runs the same steps: the NVLink barriers are collective)
        import cProfile
        import io
        import pstats

        def profiled(label, fn, n=10):
            pr = cProfile.Profile() if rank == 0 else None
            if pr is not None:
                pr.enable()
            for i in range(n):
                fn(i)
            if pr is not None:
                pr.disable()
                torch.cuda.synchronize()
                for key in ("tottime", "cumulative"):
                    buf = io.StringIO()
                    pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(40)
                    sys.stderr.write(f"==== host profile: {label} ({n} steps), sorted by {key}\n" + buf.getvalue())
            torch.cuda.synchronize()

        if rank == 0:
            torch.cuda.set_sync_debug_mode("warn")  # every host<->stream synchronisation inside a step is reported on stderr
        for i in range(2):
            step(dev_batches[i % len(dev_batches)])
        torch.cuda.set_sync_debug_mode("default")
        profiled("plain step", lambda i: step(dev_batches[i % len(dev_batches)]))
        if not args.no_e2e:
            it2 = host_iter(10 + 4)
            profiled("TrainPipelineSparseDist.progress", lambda i: pipe.progress(it2))
    if rank == 0:
        base = BASELINE_SAMPLES_PER_SEC.get(world)
        out = {
            "metric": "DLRM training throughput (samples/s, whole job, device-timed, max over ranks)",
            "value": value,
            "unit": "samples/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": (value / base) if base else None,
            "dtype": "bf16" if dense_backend == "tcgen05" else "fp32",
            "data": "synthetic (random Criteo-1TB-shaped ids/dense, random-init tables)",
            "impl": "ours",
            "config": {
                "model": "DLRM (26 sparse x dim %d, dense %s, over %s), fp32 tables + fused row-wise Adagrad, dense SGD" % (args.embedding_dim, args.dense_arch, args.over_arch),
                "num_embeddings": "Criteo-1TB cardinalities capped at %d rows (%.1f GB fp32 tables)" % (args.row_cap, sum(min(h, args.row_cap) for h in CRITEO_1TB_40M) * args.embedding_dim * 4 / 1e9),
                "global_batch": B * world,
                "per_gpu_batch": B,
                "seq_len": args.pooling,
                "parallelism": f"{args.sharding} embeddings over {world} GPU(s) + DDP dense" + (f"; tables with <= {args.dp_rows} rows data-parallel (dense SGD)" if world > 1 and args.dp_rows > 0 and args.sharding == "table_wise" else ""),
                "pipeline": "TrainPipelineSparseDist (e2e) / plain step (value)",
                "l2_policy": "inputs > L2: %d distinct batches, random rows of multi-GB tables (L2 126 MB)" % len(dev_batches),
                "dense_backend": dense_backend,
                "cuda_graphs": "dense sub-modules (fwd+bwd)" if args.cuda_graphs else "off",
                "cpu_binding": globals().get("_CPU_BINDING", "off"),
            },
            "clocks": clocks,
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_ms,
            "loss": float(loss.item()),
        }
        if e2e is not None:
            out["e2e"] = e2e
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
