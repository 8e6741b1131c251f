"""Every sharding type on one table set (reference examples/sharding tutorials): build a plan per ShardingType with the helper
generators, shard an EmbeddingBagCollection with it, run a training step and print where each shard lives.

    torchrun --nproc-per-node 2 examples/sharding_types.py          # gloo on CPU, nccl + NVLink fused paths on GPUs"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig  # noqa: E402
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection  # noqa: E402
from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward  # noqa: E402
from torchrec_b200.parallel import sharding_plan as sp  # noqa: E402
from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder  # noqa: E402
from torchrec_b200.parallel.model_parallel import DistributedModelParallel  # noqa: E402
from torchrec_b200.parallel.types import ShardingEnv, ShardingPlan  # noqa: E402
from torchrec_b200.sparse import KeyedJaggedTensor  # noqa: E402


def run(world: int, rank: int, device: torch.device) -> dict:
    results = {}
    layouts = {
        "table_wise": {"big": sp.table_wise(rank=0), "small": sp.table_wise(rank=world - 1)},
        "row_wise": {"big": sp.row_wise(), "small": sp.row_wise()},
        "column_wise": {"big": sp.column_wise(ranks=list(range(world))), "small": sp.table_wise(rank=0)},
        "data_parallel": {"big": sp.data_parallel(), "small": sp.data_parallel()},
        "mixed": {"big": sp.row_wise(), "small": sp.data_parallel()},
    }
    for name, per_table in layouts.items():
        torch.manual_seed(0)
        tables = [EmbeddingBagConfig(name="big", embedding_dim=16 * world, num_embeddings=1000, feature_names=["f_big"]),
                  EmbeddingBagConfig(name="small", embedding_dim=16 * world, num_embeddings=64, feature_names=["f_small"])]
        ebc = EmbeddingBagCollection(tables, device=torch.device("meta"))
        apply_optimizer_in_backward(torch.optim.SGD, ebc.parameters(), {"lr": 0.1})
        plan = sp.construct_module_sharding_plan(ebc, per_table, sharder=EmbeddingBagCollectionSharder(), world_size=world, local_size=world, device_type=device.type)
        model = DistributedModelParallel(ebc, env=ShardingEnv.from_process_group(dist.group.WORLD), device=device, plan=ShardingPlan({"": plan}),
                                         sharders=[EmbeddingBagCollectionSharder()])
        g = torch.Generator().manual_seed(rank)
        lengths = torch.randint(1, 4, (2 * 8,), generator=g)
        kjt = KeyedJaggedTensor(keys=["f_big", "f_small"], values=torch.randint(0, 64, (int(lengths.sum()),), generator=g), lengths=lengths).to(device)
        out = model(kjt).values()
        out.sum().backward()
        where = {t: [str(s.placement) for s in (ps.sharding_spec.shards if ps.sharding_spec is not None else [])] or ["replicated"] for t, ps in plan.items()}
        results[name] = (tuple(out.shape), where)
        if rank == 0:
            print(f"{name:14s} out {tuple(out.shape)}  " + "  ".join(f"{t}: {w}" for t, w in where.items()))
    return results


def main() -> None:
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29534")
        cuda = torch.cuda.is_available()
        dist.init_process_group("nccl" if cuda else "gloo", rank=int(os.environ.get("RANK", 0)), world_size=int(os.environ.get("WORLD_SIZE", 1)))
        created = True
    cuda = torch.cuda.is_available()
    device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}") if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    run(dist.get_world_size(), dist.get_rank(), device)
    if created:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
