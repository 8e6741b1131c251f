"""Collective micro-benchmarks (reference torchrec/distributed/benchmark/benchmark_comms.py): pooled-embedding all-to-all
through NCCL vs the fused NVLink store path, KJT all-to-all, reduce-scatter. Device-timed, max over ranks.

    torchrun --nproc-per-node 8 -m torchrec_b200.benchmarks.benchmark_comms --batch 32768 --dim 128 --features 26"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def _time(fn, iters: int = 20, warm: int = 5) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--features", type=int, default=26)
    ap.add_argument("--dtype", type=str, default="bf16")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    f_local = (a.features + world - 1) // world
    # pooled all-to-all: every rank holds [B*W, f_local*D] and receives [B, W*f_local*D]
    src = torch.randn(a.batch * world, f_local * a.dim, device=dev).to(dt)
    dst = torch.empty(world, a.batch, f_local * a.dim, device=dev, dtype=dt)
    nbytes = src.numel() * src.element_size()
    ms = _time(lambda: dist.all_to_all_single(dst.view(world * a.batch, -1), src))
    res = {"op": "pooled_all_to_all(nccl)", "ms": ms, "GBps_per_rank": nbytes * (world - 1) / world / ms / 1e6}
    if rank == 0:
        print(json.dumps(res))
    from ..parallel.p2p import PeerGroup, cast_copy

    if PeerGroup.supported(dist.group.WORLD, dev):
        pg = PeerGroup.get(dist.group.WORLD, dev)
        buf = pg.alloc(nbytes)
        views = [__import__("torchrec_b200.parallel.p2p", fromlist=["tensor_from_ptr"]).tensor_from_ptr(p, nbytes, dev).view(dt).view(world, a.batch, -1) for p in buf.ptrs]

        def p2p() -> None:
            for r in range(world):
                views[r][rank].copy_(src[r * a.batch : (r + 1) * a.batch], non_blocking=True)  # store into the destination rank's slot
            pg.barrier()

        ms = _time(p2p)
        if rank == 0:
            print(json.dumps({"op": "pooled_all_to_all(nvlink peer stores + device barrier)", "ms": ms, "GBps_per_rank": nbytes * (world - 1) / world / ms / 1e6}))
    rs_in = torch.randn(world * a.batch, a.dim, device=dev).to(dt)
    rs_out = torch.empty(a.batch, a.dim, device=dev, dtype=dt)
    ms = _time(lambda: dist.reduce_scatter_tensor(rs_out, rs_in))
    if rank == 0:
        print(json.dumps({"op": "reduce_scatter(nccl)", "ms": ms, "GBps_per_rank": rs_in.numel() * rs_in.element_size() * (world - 1) / world / ms / 1e6}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
