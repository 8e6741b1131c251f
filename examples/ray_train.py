"""Launch a training run on a Ray cluster (reference ``examples/ray/train_torchrec.py``, ``compute_world_size.py``).

Ray only does placement: one actor per GPU, each actor sets the ``torch.distributed`` environment and runs the same ``train(rank, world_size)`` function
a ``torchrun`` launch would. Ray is not part of this image, so ``main()`` falls back to local processes when it cannot be imported - the training
function is identical either way.

    python examples/ray_train.py --num-workers 2 --cpu
"""
import argparse
import os
import socket
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def train(rank: int, world_size: int, master_addr: str, master_port: int, cpu: bool = True, steps: int = 5) -> float:
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world_size), MASTER_ADDR=master_addr, MASTER_PORT=str(master_port))
    import importlib.util

    spec = importlib.util.spec_from_file_location("golden_dp", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_training_data_parallel.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.main((["--cpu"] if cpu else []) + ["--steps", str(steps)])


def compute_world_size(rank: int, world_size: int, master_addr: str, master_port: int) -> int:
    """The reference's connectivity check: all-reduce a one and compare with the world size."""
    import torch
    import torch.distributed as dist

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), MASTER_ADDR=master_addr, MASTER_PORT=str(master_port))
    dist.init_process_group("gloo")
    t = torch.ones(1)
    dist.all_reduce(t)
    dist.destroy_process_group()
    return int(t.item())


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-workers", type=int, default=2)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    port = _free_port()
    try:
        import ray  # type: ignore[import-not-found]
    except ImportError:
        import torch.multiprocessing as mp

        print("ray is not installed: launching local processes instead")
        mp.spawn(_local_entry, args=(a.num_workers, port, a.cpu, a.steps), nprocs=a.num_workers, join=True)
        return
    ray.init()
    remote_train = ray.remote(num_gpus=0 if a.cpu else 1)(train)
    addr = ray.util.get_node_ip_address()
    print(ray.get([remote_train.remote(r, a.num_workers, addr, port, a.cpu, a.steps) for r in range(a.num_workers)]))


def _local_entry(rank: int, world: int, port: int, cpu: bool, steps: int) -> None:
    train(rank, world, "127.0.0.1", port, cpu, steps)


if __name__ == "__main__":
    main()
