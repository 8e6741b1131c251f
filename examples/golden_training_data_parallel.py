"""The same DLRM trained purely data-parallel: every table replicated, dense gradients all-reduced by DDP
(reference ``examples/golden_training/train_dlrm_data_parallel.py``). Useful as the baseline a model-parallel plan is compared against: identical model,
identical data, only the plan differs (``data_parallel()`` for every table, dense compute kernel).

    torchrun --nproc-per-node 8 examples/golden_training_data_parallel.py
    python examples/golden_training_data_parallel.py --cpu
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchrec_b200.datasets.random import RandomRecDataset  # noqa: E402
from torchrec_b200.models.dlrm import DLRM, DLRMTrain  # noqa: E402
from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig  # noqa: E402
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection  # noqa: E402
from torchrec_b200.optim.keyed import KeyedOptimizerWrapper  # noqa: E402
from torchrec_b200.parallel import sharding_plan as sp  # noqa: E402
from torchrec_b200.parallel.model_parallel import DistributedModelParallel  # noqa: E402
from torchrec_b200.parallel.types import ShardingPlan  # noqa: E402


def main(argv=None) -> float:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch-size", type=int, default=256)
    a = ap.parse_args(argv)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    cuda = torch.cuda.is_available() and not a.cpu
    device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}") if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if cuda else "gloo")
    torch.manual_seed(0)
    keys = [f"cat_{i}" for i in range(8)]
    hashes = [5000] * 8
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name=f"t_{k}", embedding_dim=16, num_embeddings=h, feature_names=[k]) for k, h in zip(keys, hashes)], device=device)
    train_model = DLRMTrain(DLRM(ebc, 13, [32, 16], [32, 1], dense_device=device))
    module_plan = sp.construct_module_sharding_plan(ebc, {f"t_{k}": sp.data_parallel() for k in keys}, world_size=world, local_size=world, device_type=device.type)
    model = DistributedModelParallel(train_model, device=device, plan=ShardingPlan({"model.sparse_arch.embedding_bag_collection": module_plan}))
    opt = KeyedOptimizerWrapper(dict(model.named_parameters()), lambda p: torch.optim.Adagrad(p, lr=0.05))
    data = iter(RandomRecDataset(keys, a.batch_size, hash_sizes=hashes, ids_per_features=[2] * 8, num_dense=13, manual_seed=rank, num_batches=a.steps + 1))
    loss = torch.zeros(())
    for step in range(a.steps):
        batch = next(data).to(device)
        loss, _ = model(batch)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if rank == 0 and step % 5 == 0:
            print(f"step {step}: loss {float(loss):.4f}")
    return float(loss)


if __name__ == "__main__":
    main()
    if dist.is_initialized():
        dist.destroy_process_group()
