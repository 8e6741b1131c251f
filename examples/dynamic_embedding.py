"""Dynamic embedding (reference contrib/dynamic_embedding): train over an unbounded id space with a fixed-size cache table;
rows live in a parameter server (here: the built-in file:// backend) and move through the native id transformer."""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchrec_b200.dynamic_embedding import wrap  # noqa: E402
from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig  # noqa: E402
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection  # noqa: E402
from torchrec_b200.sparse import KeyedJaggedTensor  # noqa: E402


class Batch:
    def __init__(self, kjt, y):
        self.sparse_features, self.labels = kjt, y


class Model(torch.nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="t", embedding_dim=8, num_embeddings=256, feature_names=["f"])])  # 256 = CACHE rows
        self.lin = torch.nn.Linear(8, 1)

    def forward(self, kjt):
        return self.lin(self.ebc(kjt).values()).squeeze(-1)


def main() -> None:
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)
    data = []
    for _ in range(40):
        ids = torch.randint(0, 2**50, (32,), generator=g) % 5000 * 1_000_003  # 5000 distinct huge ids >> 256 cache rows
        data.append(Batch(KeyedJaggedTensor(keys=["f"], values=ids, lengths=torch.ones(32, dtype=torch.int64)), (ids % 2).float()))
    model = Model()
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    with tempfile.TemporaryDirectory() as d:
        loader, colls = wrap(f"file://{d}/ps", data, model, eviction_config={"type": "mixed_lru_lfu"})
        for step, batch in enumerate(loader):
            loss = torch.nn.functional.binary_cross_entropy_with_logits(model(batch.sparse_features), batch.labels)
            opt.zero_grad()
            loss.backward()
            opt.step()
            if step % 10 == 0:
                print(f"step {step}: loss {float(loss):.4f}  cached ids {len(colls[0].transformers[0])}")
        colls[0].save()
        print("rows in the parameter server:", sum(len(p) for p in colls[0]._ps.values()))


if __name__ == "__main__":
    main()
