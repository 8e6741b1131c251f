"""Golden DLRM training loop (reference examples/golden_training/train_dlrm.py). One process per GPU:

    torchrun --nproc-per-node 8 examples/golden_training.py            # NCCL + NVLink fused paths
    python examples/golden_training.py --cpu                           # single process smoke run
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
from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward  # noqa: E402
from torchrec_b200.optim.keyed import CombinedOptimizer, KeyedOptimizerWrapper  # noqa: E402
from torchrec_b200.optim.optimizers import in_backward_optimizer_filter  # noqa: E402
from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad  # noqa: E402
from torchrec_b200.parallel.model_parallel import DistributedModelParallel  # noqa: E402
from torchrec_b200.parallel.train_pipeline import TrainPipelineSparseDist  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch-size", type=int, default=1024)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cuda = torch.cuda.is_available() and not a.cpu
    device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}") if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if cuda else "gloo")
    keys = [f"cat_{i}" for i in range(26)]
    hashes = [100_000] * 26
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name=f"t_{k}", embedding_dim=64, num_embeddings=h, feature_names=[k]) for k, h in zip(keys, hashes)], device=torch.device("meta"))
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.02})
    train_model = DLRMTrain(DLRM(ebc, 13, [128, 64], [256, 128, 1], dense_device=device))
    model = DistributedModelParallel(train_model, device=device)  # plan=None -> the planner shards the tables (collectively)
    dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(model.named_parameters())), lambda p: torch.optim.Adagrad(p, lr=0.02))
    optimizer = CombinedOptimizer([model.fused_optimizer, dense_opt])
    data = iter(RandomRecDataset(keys, a.batch_size, hash_sizes=hashes, ids_per_features=[3] * 26, num_dense=13, manual_seed=rank, num_batches=a.steps + 4, pin_memory=cuda))
    pipeline = TrainPipelineSparseDist(model, optimizer, device)
    for step in range(a.steps):
        loss, logits, labels = pipeline.progress(data)  # DLRMTrain returns (loss, (loss, logits, labels)); the pipeline hands back the tuple
        if rank == 0 and step % 5 == 0:
            print(f"step {step}: loss {float(loss):.4f}")
    if rank == 0:
        print("plan:\n", model.plan)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
