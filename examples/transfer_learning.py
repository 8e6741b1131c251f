"""Transfer learning (reference examples/transfer_learning): start an EmbeddingBagCollection from a pretrained embedding matrix
(e.g. word / item vectors trained elsewhere), freeze or fine-tune it, and check that the sharded model serves the same vectors.

    python examples/transfer_learning.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig  # noqa: E402
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection  # noqa: E402
from torchrec_b200.sparse import KeyedJaggedTensor  # noqa: E402


def load_pretrained(ebc: EmbeddingBagCollection, pretrained: dict, freeze: bool = False) -> None:
    """Copy ``{table_name: [rows, dim] tensor}`` into the collection through its state dict (the same keys a sharded model uses, so
    the call works unchanged on ``DistributedModelParallel(ebc)``: full tensors are sliced per shard on load)."""
    sd = {f"embedding_bags.{name}.weight": w for name, w in pretrained.items()}
    missing, unexpected = ebc.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    if freeze:
        for name in pretrained:
            ebc.embedding_bags[name].weight.requires_grad_(False)


def main() -> float:
    torch.manual_seed(0)
    rows, dim = 500, 16
    pretrained = {"item": torch.randn(rows, dim) * 0.3}
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="item", embedding_dim=dim, num_embeddings=rows, feature_names=["item_id"]),
                                  EmbeddingBagConfig(name="user", embedding_dim=dim, num_embeddings=100, feature_names=["user_id"])])
    load_pretrained(ebc, pretrained, freeze=True)
    kjt = KeyedJaggedTensor(keys=["item_id", "user_id"], values=torch.tensor([3, 7, 9, 1, 2]), lengths=torch.tensor([2, 1, 1, 1]))
    out = ebc(kjt).to_dict()
    torch.testing.assert_close(out["item_id"][0], pretrained["item"][3] + pretrained["item"][7])  # pooled pretrained vectors
    # fine-tune only the user table against the frozen item space
    head = torch.nn.Linear(2 * dim, 1)
    opt = torch.optim.SGD([p for p in list(ebc.parameters()) + list(head.parameters()) if p.requires_grad], lr=0.05)
    first = last = 0.0
    for step in range(60):
        pooled = ebc(kjt).values()
        loss = (head(pooled).squeeze(1) - torch.tensor([1.0, 0.0])).pow(2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        first = float(loss.detach()) if step == 0 else first
        last = float(loss.detach())
    torch.testing.assert_close(ebc.embedding_bags["item"].weight.detach(), pretrained["item"])  # frozen table untouched
    print(f"fine-tuned on frozen item embeddings: loss {first:.4f} -> {last:.4f}")
    return last


if __name__ == "__main__":
    main()
