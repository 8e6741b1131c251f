"""Zero-collision hashing (reference examples/zch): raw 64-bit ids -> managed-collision remap -> embedding lookup.
Runs single process; under torchrun the same module is sharded by ManagedCollisionEmbeddingBagCollectionSharder."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig  # noqa: E402
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection  # noqa: E402
from torchrec_b200.modules.hash_mc_modules import HashZchManagedCollisionModule  # noqa: E402
from torchrec_b200.modules.mc_embedding_modules import ManagedCollisionEmbeddingBagCollection  # noqa: E402
from torchrec_b200.modules.mc_modules import DistanceLFU_EvictionPolicy, ManagedCollisionCollection, MCHManagedCollisionModule  # noqa: E402
from torchrec_b200.sparse import KeyedJaggedTensor  # noqa: E402


def main() -> None:
    torch.manual_seed(0)
    dev = torch.device("cpu")
    tables = [EmbeddingBagConfig(name="t_user", embedding_dim=16, num_embeddings=4096, feature_names=["user"]),
              EmbeddingBagConfig(name="t_item", embedding_dim=16, num_embeddings=4096, feature_names=["item"])]
    mcs = {"t_user": MCHManagedCollisionModule(zch_size=4096, device=dev, eviction_policy=DistanceLFU_EvictionPolicy(), eviction_interval=4, input_hash_size=2**62),
           "t_item": HashZchManagedCollisionModule(zch_size=4096, device=dev, total_num_buckets=4, max_probe=64)}  # sorted ZCH and multi-probe ZCH side by side
    model = ManagedCollisionEmbeddingBagCollection(EmbeddingBagCollection(tables, device=dev), ManagedCollisionCollection(mcs, tables), return_remapped_features=True)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(0)
    universe = torch.randint(0, 2**60, (3000,), generator=g)
    for step in range(12):
        ids = universe[torch.randint(0, 3000, (64,), generator=g)]
        kjt = KeyedJaggedTensor(keys=["user", "item"], values=torch.cat([ids[:32], ids[32:]]), lengths=torch.ones(64, dtype=torch.int64))
        pooled, remapped = model(kjt)
        loss = pooled.values().pow(2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 4 == 0:
            print(f"step {step}: loss {float(loss):.5f}  max slot {int(remapped.values().max())}  open slots {[int(v) for v in model._managed_collision_collection.open_slots().values()]}")


if __name__ == "__main__":
    main()
