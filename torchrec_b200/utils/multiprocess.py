"""Local multi-process test harness: spawn ``world_size`` processes on one host with a free port
(same methodology as the reference's MultiProcessTestBase, distributed/test_utils/multi_process.py)."""
from __future__ import annotations

import multiprocessing
import os
import socket
import traceback
from typing import Any, Callable, Dict, Optional

import torch
import torch.distributed as dist


def get_free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class MultiProcessContext:
    def __init__(self, rank: int, world_size: int, backend: str = "gloo", local_size: Optional[int] = None, use_cuda: Optional[bool] = None) -> None:
        self.rank = rank
        self.world_size = world_size
        self.backend = backend
        self.local_size = local_size
        if use_cuda is None:
            use_cuda = backend == "nccl"
        if use_cuda and torch.cuda.is_available() and torch.cuda.device_count() >= world_size:
            self.device = torch.device(f"cuda:{rank}")
            torch.cuda.set_device(self.device)
        else:
            self.device = torch.device("cpu")
        self.pg: Optional[dist.ProcessGroup] = None

    def __enter__(self) -> "MultiProcessContext":
        os.environ["RANK"] = str(self.rank)
        os.environ["WORLD_SIZE"] = str(self.world_size)
        os.environ["LOCAL_RANK"] = str(self.rank if self.local_size is None else self.rank % self.local_size)
        if self.local_size is not None:
            os.environ["LOCAL_WORLD_SIZE"] = str(self.local_size)
        else:
            os.environ["LOCAL_WORLD_SIZE"] = str(self.world_size)
        dist.init_process_group(backend=self.backend, rank=self.rank, world_size=self.world_size)
        self.pg = dist.group.WORLD
        return self

    def __exit__(self, *exc) -> None:
        if dist.is_initialized():
            try:
                dist.barrier()
            except Exception:
                pass
            dist.destroy_process_group()


def _entry(rank: int, world_size: int, fn: Callable, port: int, backend: str, local_size: Optional[int], kwargs: Dict[str, Any], err_q) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["GLOO_DEVICE_TRANSPORT"] = "TCP"
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    os.environ.setdefault("NCCL_SOCKET_IFNAME", "lo")
    torch.set_num_threads(1)
    import faulthandler

    faulthandler.enable()  # a native crash (SIGSEGV in a kernel launcher, NCCL, IPC) prints the Python stack of the rank
    try:
        with MultiProcessContext(rank, world_size, backend, local_size) as ctx:
            fn(ctx, **kwargs)
    except Exception:
        err_q.put((rank, traceback.format_exc()))
        raise


def run_multi_process(fn: Callable, world_size: int = 2, backend: str = "gloo", local_size: Optional[int] = None, timeout: float = 240.0, **kwargs: Any) -> None:
    """Run ``fn(ctx, **kwargs)`` in ``world_size`` spawned processes; raises if any rank fails."""
    mp = multiprocessing.get_context("spawn")
    port = get_free_port()
    err_q = mp.Queue()
    procs = []
    for r in range(world_size):
        p = mp.Process(target=_entry, args=(r, world_size, fn, port, backend, local_size, kwargs, err_q))
        p.start()
        procs.append(p)
    failed = []
    for p in procs:
        p.join(timeout)
        if p.is_alive():
            p.terminate()
            failed.append("timeout")
        elif p.exitcode != 0:
            failed.append(f"exit {p.exitcode}")
    if failed:
        msgs = []
        while not err_q.empty():
            r, tb = err_q.get()
            msgs.append(f"--- rank {r} ---\n{tb}")
        raise RuntimeError(f"multi-process run failed: {failed}\n" + "\n".join(msgs))
